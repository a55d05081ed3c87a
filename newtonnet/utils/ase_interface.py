"""newtonnet/utils/ase_interface.py of the reference -> newtonnet_b200.utils.ase_interface."""
from newtonnet_b200.utils.ase_interface import *             # noqa: F401,F403
from newtonnet_b200.utils.ase_interface import __all__       # noqa: F401
