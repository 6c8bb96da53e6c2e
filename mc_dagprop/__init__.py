"""``mc_dagprop`` -- alias package: the reference's import path on the B200 engine.

``from mc_dagprop import Simulator, DagContext, ...`` keeps working for code written against
WonJayne/mc_dagprop (reference ``src/mc_dagprop/__init__.py``); every name resolves to
``mc_dagprop_b200``.  The analytic (PMF) propagator of the reference is a different algorithm and
is not part of this package (out of scope).
"""
from mc_dagprop_b200 import __version__
from mc_dagprop_b200.monte_carlo import (
    Activity,
    DagContext,
    Event,
    EventTimestamp,
    GenericDelayGenerator,
    MonteCarloPropagator,
    SimResult,
    Simulator,
)

__all__ = [
    "GenericDelayGenerator",
    "DagContext",
    "SimResult",
    "Event",
    "Activity",
    "Simulator",
    "MonteCarloPropagator",
    "EventTimestamp",
    "__version__",
]
