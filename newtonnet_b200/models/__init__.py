from newtonnet_b200.models.newtonnet import *  # noqa: F401,F403
from newtonnet_b200.models.output import *  # noqa: F401,F403
