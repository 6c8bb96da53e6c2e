from mc_dagprop_b200.analytic._propagator import *  # noqa: F401,F403
