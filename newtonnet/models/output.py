"""newtonnet/models/output.py of the reference -> newtonnet_b200.models.output."""
from newtonnet_b200.models.output import *                 # noqa: F401,F403
from newtonnet_b200.models.output import __all__           # noqa: F401
