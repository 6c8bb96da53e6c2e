from mc_dagprop_b200.types import *  # noqa: F401,F403
from mc_dagprop_b200.types import __all__  # noqa: F401
