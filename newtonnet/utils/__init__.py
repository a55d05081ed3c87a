from newtonnet.utils.pretrained_models import *  # noqa: F401,F403
from newtonnet.utils.ase_interface import *      # noqa: F401,F403
