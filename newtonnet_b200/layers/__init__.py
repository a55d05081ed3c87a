from newtonnet_b200.layers.activations import *  # noqa: F401,F403
from newtonnet_b200.layers.precision import *  # noqa: F401,F403
from newtonnet_b200.layers.representations import *  # noqa: F401,F403
from newtonnet_b200.layers.scalers import *  # noqa: F401,F403
