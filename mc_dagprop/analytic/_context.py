from mc_dagprop_b200.analytic._context import *  # noqa: F401,F403
