"""newtonnet/layers/activations.py of the reference -> newtonnet_b200.layers.activations."""
from newtonnet_b200.layers.activations import *            # noqa: F401,F403
from newtonnet_b200.layers.activations import __all__      # noqa: F401
