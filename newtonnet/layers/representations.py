"""newtonnet/layers/representations.py of the reference -> newtonnet_b200.layers.representations."""
from newtonnet_b200.layers.representations import *            # noqa: F401,F403
from newtonnet_b200.layers.representations import __all__      # noqa: F401
