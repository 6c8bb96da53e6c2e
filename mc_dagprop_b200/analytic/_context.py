"""Containers of the analytic propagator (reference ``analytic/_context.py:14-90``); the validation of a context is
in ``_validate.py``."""
from __future__ import annotations

from dataclasses import dataclass
from enum import IntEnum, unique

from ..monte_carlo import Event
from ..types import ActivityIndex, EventIndex, ProbabilityMass, Second
from ._pmf import DiscretePMF

PredecessorTuple = tuple[EventIndex, ActivityIndex]


@dataclass(frozen=True, slots=True)
class AnalyticActivity:
    """An edge and the distribution of its delay."""

    idx: ActivityIndex
    pmf: DiscretePMF


@dataclass(frozen=True, slots=True)
class SimulatedEvent:
    """Distribution of an event's time plus the mass that fell below / above its bounds."""

    pmf: DiscretePMF
    underflow: ProbabilityMass
    overflow: ProbabilityMass


@unique
class UnderflowRule(IntEnum):
    """Mass below the lower bound: moved onto the bound, dropped, or spread over the kept bins."""

    TRUNCATE = 1
    REMOVE = 2
    REDISTRIBUTE = 3


@unique
class OverflowRule(IntEnum):
    """Mass above the upper bound: moved onto the bound, dropped, or spread over the kept bins."""

    TRUNCATE = 1
    REMOVE = 2
    REDISTRIBUTE = 3


@dataclass(frozen=True, slots=True)
class AnalyticContext:
    """The network: events, ``(src, dst) -> (activity index, AnalyticActivity)``, precedence list, grid ``step``,
    flow rules and the optional cap ``upper = min(latest, earliest + max_delay)``."""

    events: tuple[Event, ...]
    activities: dict[tuple[EventIndex, EventIndex], tuple[ActivityIndex, AnalyticActivity]]
    precedence_list: tuple[tuple[EventIndex, tuple[PredecessorTuple, ...]], ...]
    step: Second
    underflow_rule: UnderflowRule
    overflow_rule: OverflowRule
    max_delay: Second | None = None


from ._validate import validate_context  # noqa: E402,F401  (the checks live in their own module)
