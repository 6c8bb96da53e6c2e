from mc_dagprop_b200.monte_carlo import *  # noqa: F401,F403
from mc_dagprop_b200.monte_carlo import __all__, _core  # noqa: F401
