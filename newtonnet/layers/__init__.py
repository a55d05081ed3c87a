from newtonnet.layers.activations import *       # noqa: F401,F403
from newtonnet.layers.precision import *         # noqa: F401,F403
from newtonnet.layers.representations import *   # noqa: F401,F403
from newtonnet.layers.scalers import *           # noqa: F401,F403
