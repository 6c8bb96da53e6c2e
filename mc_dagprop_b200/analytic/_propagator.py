"""``AnalyticPropagator``: deterministic propagation of discrete PMFs through the DAG on the GPU
(reference ``analytic/_propagator.py``).

``run()`` flattens the context into arrays -- integer bounds of every event, the precedence list, one PMF per
edge -- and makes ONE call into the library (``mcdp_analytic_run``): a kernel launch per topological level, a CTA per
event, convolution / maximum / bound handling block-cooperative in double-double arithmetic.  The result comes back
as one packed array of bins that is sliced into the reference's ``SimulatedEvent`` objects.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..types import ActivityIndex, EventIndex, ProbabilityMass, Second
from . import _device
from ._context import AnalyticContext, OverflowRule, PredecessorTuple, SimulatedEvent, UnderflowRule, validate_context
from ._pmf import DiscretePMF


def _topology(context: AnalyticContext):
    """Predecessors by target and a topological order (Kahn, FIFO over ascending roots: the reference's order)."""
    n = len(context.events)
    preds_of: list[tuple[PredecessorTuple, ...] | None] = [None] * n
    successors: list[list[int]] = [[] for _ in range(n)]
    indegree = [0] * n
    for target, preds in context.precedence_list:
        preds_of[target] = preds
        indegree[target] = len(preds)
        for src, _ in preds:
            successors[src].append(target)
    order = [i for i, deg in enumerate(indegree) if deg == 0]
    head = 0
    while head < len(order):
        node = order[head]
        head += 1
        for dst in successors[node]:
            indegree[dst] -= 1
            if indegree[dst] == 0:
                order.append(dst)
    if len(order) != n:
        raise RuntimeError("Invalid DAG: cycle detected")
    return tuple(preds_of), tuple(order)


def create_analytic_propagator(context: AnalyticContext, validate: bool = True) -> "AnalyticPropagator":
    """Build the propagator for ``context``; ``validate=False`` skips :func:`validate_context`."""
    if validate:
        validate_context(context)
    preds_of, order = _topology(context)
    return AnalyticPropagator(context=context, _predecessors_by_target=preds_of, _topological_node_order=order)


@dataclass(frozen=True, slots=True)
class AnalyticPropagator:
    """Propagates PMFs through the DAG.  Mass outside an event's bounds is handled by the context's rules."""

    context: AnalyticContext
    _predecessors_by_target: tuple[tuple[PredecessorTuple, ...] | None, ...]
    _topological_node_order: tuple[EventIndex, ...]

    @property
    def underflow_rule(self) -> UnderflowRule:
        return self.context.underflow_rule

    @property
    def overflow_rule(self) -> OverflowRule:
        return self.context.overflow_rule

    def _event_bounds(self, earliest: Second, latest: Second) -> tuple[int, int]:
        upper = latest
        if self.context.max_delay is not None:
            upper = min(latest, earliest + self.context.max_delay)
        return int(np.round(earliest)), int(np.round(upper))

    def run(self) -> tuple[SimulatedEvent, ...]:
        """One ``SimulatedEvent`` per event, in the order of ``context.events``."""
        ctx = self.context
        step = ctx.step
        n = len(ctx.events)
        lower, upper, origin = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
        for i, ev in enumerate(ctx.events):
            lower[i], upper[i] = self._event_bounds(ev.timestamp.earliest, ev.timestamp.latest)
            origin[i] = int(round(round(ev.timestamp.earliest / step) * step))
        # one PMF per distinct activity object, in first-use order
        pmf_index: dict[int, int] = {}
        pmf_start: list[int] = []
        pmf_probs: list[np.ndarray] = []
        targets, offsets, srcs, pmfs = [], [0], [], []
        for target, preds in enumerate(self._predecessors_by_target):
            if not preds:
                continue
            targets.append(target)
            for src, _ in preds:
                activity = ctx.activities[(src, target)][1]
                key = id(activity.pmf)
                if key not in pmf_index:
                    pmf_index[key] = len(pmf_start)
                    pmf_start.append(int(round(float(activity.pmf.values[0]))))
                    pmf_probs.append(np.ascontiguousarray(activity.pmf.probabilities, np.float64))
                srcs.append(src)
                pmfs.append(pmf_index[key])
            offsets.append(len(srcs))
        pmf_off = np.concatenate([[0], np.cumsum([p.size for p in pmf_probs])]).astype(np.int64)
        flat = np.concatenate(pmf_probs) if pmf_probs else np.zeros(0)
        start, length, off, probs, under, over = _device.analytic_run(
            lower, upper, origin, int(step), targets, offsets, srcs, pmfs, pmf_start, pmf_off, flat,
            int(self.underflow_rule), int(self.overflow_rule))
        out = []
        for i in range(n):
            p = probs[off[i]: off[i] + length[i]].copy()
            values = float(start[i]) + float(step) * np.arange(p.size, dtype=float)
            out.append(SimulatedEvent(DiscretePMF(values, p, step=step), ProbabilityMass(under[i]), ProbabilityMass(over[i])))
        return tuple(out)

    def _convert_to_simulated_event(self, pmf: DiscretePMF, min_value: int, max_value: int) -> SimulatedEvent:
        """Clip ``pmf`` to ``[min_value, max_value]`` under the flow rules (reference ``_propagator.py:158-265``)."""
        if min_value > max_value:
            raise ValueError("min_value must not exceed max_value")
        if pmf.values.size == 0:
            raise ValueError("PMF must not be empty")
        start, probs, under, over = _device.pmf_op(
            2, pmf.step, int(round(float(pmf.values[0]))), pmf.probabilities, bounds=(int(np.round(min_value)), int(np.round(max_value))),
            rules=(int(self.underflow_rule), int(self.overflow_rule)))
        values = float(start) + float(pmf.step) * np.arange(probs.size, dtype=float)
        return SimulatedEvent(DiscretePMF(values, probs, step=pmf.step), ProbabilityMass(under), ProbabilityMass(over))
