"""newtonnet/models/newtonnet.py of the reference -> newtonnet_b200.models.newtonnet."""
from newtonnet_b200.models.newtonnet import *              # noqa: F401,F403
from newtonnet_b200.models.newtonnet import EmbeddingNet, InteractionNet, NewtonNet, __all__    # noqa: F401
