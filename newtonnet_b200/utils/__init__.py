from newtonnet_b200.utils.pretrained_models import *  # noqa: F401,F403
from newtonnet_b200.utils.ase_interface import *  # noqa: F401,F403
