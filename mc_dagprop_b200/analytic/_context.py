"""Containers of the analytic propagator and their validation (reference ``analytic/_context.py``)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import IntEnum, unique

import numpy as np

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


def validate_context(context: AnalyticContext) -> None:
    """Structural checks of the reference (``_context.py:93-158``): step, ``max_delay``, event windows, activity
    indices / alignment / unit mass, precedence indices and acyclicity.  Every violation is a ``ValueError``."""
    n = len(context.events)
    if context.step <= 0.0:
        raise ValueError("step_size must be positive")
    if context.max_delay is not None and context.max_delay < 0.0:
        raise ValueError("max_delay must be non-negative when provided")
    for i, ev in enumerate(context.events):
        ts = ev.timestamp
        if ts.earliest > ts.latest:
            raise ValueError(f"event {i} has earliest > latest")
        if not (ts.earliest <= ts.actual <= ts.latest):
            raise ValueError(f"event {i} actual time outside bounds")
    for (src, dst), (_, edge) in context.activities.items():
        if not (0 <= src < n and 0 <= dst < n):
            raise ValueError(f"activity {(src, dst)} references invalid node")
        edge.pmf.validate()
        if not np.isclose(edge.pmf.step, context.step):
            raise ValueError(f"edge {(src, dst)} step {edge.pmf.step} does not match context step size {context.step}")
        edge.pmf.validate_alignment(context.step)
        if not np.isclose(edge.pmf.total_mass, 1.0):
            raise ValueError(f"activity {(src, dst)} PMF does not sum to 1, got {edge.pmf.total_mass}")
    indegree = [0] * n
    successors: list[list[int]] = [[] for _ in range(n)]
    for target, preds in context.precedence_list:
        if not (0 <= target < n):
            raise ValueError(f"target index {target} out of range")
        for src, link in preds:
            if not (0 <= src < n):
                raise ValueError(f"predecessor index {src} out of range")
            edge = context.activities.get((src, target))
            if edge is None:
                raise ValueError(f"missing activity for {(src, target)}")
            if edge[0] != link:
                raise ValueError(f"edge index {link} for {(src, target)} does not match context mapping {edge[0]}")
            successors[src].append(target)
            indegree[target] += 1
    ready = [i for i, deg in enumerate(indegree) if deg == 0]
    seen = 0
    while ready:
        node = ready.pop()
        seen += 1
        for dst in successors[node]:
            indegree[dst] -= 1
            if indegree[dst] == 0:
                ready.append(dst)
    if seen != n:
        raise ValueError("precedence list contains a cycle")
