from mc_dagprop_b200.analytic.distributions import *  # noqa: F401,F403
