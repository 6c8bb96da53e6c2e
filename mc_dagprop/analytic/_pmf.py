from mc_dagprop_b200.analytic._pmf import *  # noqa: F401,F403
