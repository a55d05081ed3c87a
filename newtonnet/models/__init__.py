from newtonnet.models.newtonnet import *    # noqa: F401,F403
from newtonnet.models.output import *       # noqa: F401,F403
