from mc_dagprop_b200.analytic import *  # noqa: F401,F403
from mc_dagprop_b200.analytic import __all__, _context, _pmf, _propagator, distributions  # noqa: F401
