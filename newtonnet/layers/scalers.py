"""newtonnet/layers/scalers.py of the reference -> newtonnet_b200.layers.scalers."""
from newtonnet_b200.layers.scalers import *            # noqa: F401,F403
from newtonnet_b200.layers.scalers import __all__      # noqa: F401
