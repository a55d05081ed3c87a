"""newtonnet/layers/precision.py of the reference -> newtonnet_b200.layers.precision."""
from newtonnet_b200.layers.precision import *            # noqa: F401,F403
from newtonnet_b200.layers.precision import __all__      # noqa: F401
