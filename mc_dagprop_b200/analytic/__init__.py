"""Analytic (PMF) propagation on B200.

The second engine of the package: instead of sampling, every event's arrival time is carried as a probability mass
function on the integer grid of ``step`` seconds; an event takes the maximum over its predecessors of
``PMF(predecessor) (*) PMF(activity)`` and is clipped to its ``[earliest, latest]`` window under the context's
underflow / overflow rules.  The names exported here are those of the reference's ``mc_dagprop.analytic``
(``analytic/__init__.py``); the arithmetic behind them runs on the GPU (``csrc/mcdp_analytic.cu``):

======================================  ====================================================================
``DiscretePMF.convolve`` / ``maximum``   one device call each (``mcdp_pmf_op``), compensated (twice-working-precision) accumulation
``AnalyticPropagator.run``               ONE device call for the whole DAG (``mcdp_analytic_run``): a launch per
                                         topological level, a CTA per event
``create_analytic_propagator``           host side: validation (``_validate.py``) and topological order
``exponential_pmf`` ...                  host side: discretised delay distributions (``distributions.py``)
======================================  ====================================================================

Additive: ``set_device(i)`` / ``get_device()`` choose the CUDA device the engine runs on (default 0).  Like the
Monte-Carlo path there is no CPU execution path: without a CUDA device the device calls raise ``RuntimeError``.

    >>> from mc_dagprop.analytic import AnalyticActivity, AnalyticContext, DiscretePMF, OverflowRule, UnderflowRule
    >>> from mc_dagprop import Event, EventTimestamp, create_analytic_propagator
    >>> ctx = AnalyticContext(events=(Event("a", EventTimestamp(0.0, 0.0, 0.0)), Event("b", EventTimestamp(0.0, 9.0, 0.0))),
    ...                       activities={(0, 1): (0, AnalyticActivity(0, DiscretePMF(np.array([1., 2.]), np.array([.5, .5]), step=1)))},
    ...                       precedence_list=((1, ((0, 0),)),), step=1,
    ...                       underflow_rule=UnderflowRule.TRUNCATE, overflow_rule=OverflowRule.TRUNCATE)
    >>> [e.pmf.probabilities for e in create_analytic_propagator(ctx).run()]      # doctest: +SKIP
    [array([1.]), array([0.5, 0.5])]
"""
from __future__ import annotations

from ..types import ActivityIndex, EventIndex, ProbabilityMass, Second  # noqa: F401
from ._context import AnalyticActivity, AnalyticContext, OverflowRule, SimulatedEvent, UnderflowRule  # noqa: F401
from ._device import get_device, set_device  # noqa: F401
from ._pmf import DiscretePMF  # noqa: F401
from ._propagator import AnalyticPropagator, create_analytic_propagator  # noqa: F401
from .distributions import constant_pmf, empirical_pmf, exponential_pmf, gamma_pmf  # noqa: F401

_CLASSES = ("DiscretePMF", "SimulatedEvent", "UnderflowRule", "OverflowRule", "AnalyticContext", "AnalyticPropagator", "AnalyticActivity")
_FUNCTIONS = ("create_analytic_propagator", "exponential_pmf", "gamma_pmf", "constant_pmf", "empirical_pmf")
_TYPES = ("Second", "ProbabilityMass", "EventIndex", "ActivityIndex")
__all__ = [*_CLASSES, *_FUNCTIONS, *_TYPES]
