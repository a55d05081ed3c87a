"""newtonnet/utils/pretrained_models.py of the reference -> newtonnet_b200.utils.pretrained_models."""
from newtonnet_b200.utils.pretrained_models import *             # noqa: F401,F403
from newtonnet_b200.utils.pretrained_models import __all__       # noqa: F401
