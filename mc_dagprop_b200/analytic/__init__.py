"""Analytic (PMF) propagation on B200 -- the reference's ``mc_dagprop.analytic`` surface (``analytic/__init__.py``)
on top of ``csrc/mcdp_analytic.cu``."""
from __future__ import annotations

from ..types import ActivityIndex, EventIndex, ProbabilityMass, Second
from ._context import AnalyticActivity, AnalyticContext, OverflowRule, SimulatedEvent, UnderflowRule
from ._device import get_device, set_device
from ._pmf import DiscretePMF
from ._propagator import AnalyticPropagator, create_analytic_propagator
from .distributions import constant_pmf, empirical_pmf, exponential_pmf, gamma_pmf

__all__ = [
    "DiscretePMF",
    "SimulatedEvent",
    "UnderflowRule",
    "OverflowRule",
    "AnalyticContext",
    "AnalyticPropagator",
    "AnalyticActivity",
    "create_analytic_propagator",
    "exponential_pmf",
    "gamma_pmf",
    "constant_pmf",
    "empirical_pmf",
    "Second",
    "ProbabilityMass",
    "EventIndex",
    "ActivityIndex",
]
