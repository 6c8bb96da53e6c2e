"""``mc_dagprop`` -- alias package: the reference's import path on the B200 engine.

``from mc_dagprop import Simulator, DagContext, ...`` keeps working for code written against
WonJayne/mc_dagprop (reference ``src/mc_dagprop/__init__.py``); every name resolves to
``mc_dagprop_b200`` -- the Monte-Carlo classes and the analytic (PMF) propagator alike.
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

from mc_dagprop_b200.analytic import (  # noqa: E402
    AnalyticContext,
    AnalyticPropagator,
    DiscretePMF,
    OverflowRule,
    SimulatedEvent,
    UnderflowRule,
    create_analytic_propagator,
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
    "DiscretePMF",
    "SimulatedEvent",
    "UnderflowRule",
    "OverflowRule",
    "AnalyticContext",
    "AnalyticPropagator",
    "create_analytic_propagator",
    "__version__",
]
