"""``mc_dagprop`` -- alias package: the reference's import path on the B200 engine.

``from mc_dagprop import Simulator, DagContext, ...`` keeps working for code written against
WonJayne/mc_dagprop (reference ``src/mc_dagprop/__init__.py``); every name resolves to
``mc_dagprop_b200`` -- the Monte-Carlo classes and the analytic (PMF) propagator alike.
"""
from mc_dagprop_b200 import __version__
from mc_dagprop_b200.analytic import AnalyticContext, AnalyticPropagator, DiscretePMF, OverflowRule, SimulatedEvent  # noqa: F401
from mc_dagprop_b200.analytic import UnderflowRule, create_analytic_propagator  # noqa: F401
from mc_dagprop_b200.monte_carlo import Activity, DagContext, Event, EventTimestamp, GenericDelayGenerator  # noqa: F401
from mc_dagprop_b200.monte_carlo import MonteCarloPropagator, SimResult, Simulator  # noqa: F401

_MONTE_CARLO = ("GenericDelayGenerator", "DagContext", "SimResult", "Event", "Activity", "Simulator", "MonteCarloPropagator", "EventTimestamp")
_ANALYTIC = ("DiscretePMF", "SimulatedEvent", "UnderflowRule", "OverflowRule", "AnalyticContext", "AnalyticPropagator", "create_analytic_propagator")
__all__ = [*_MONTE_CARLO, *_ANALYTIC, "__version__"]
