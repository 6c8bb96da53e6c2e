"""``AnalyticPropagator``: deterministic propagation of discrete PMFs through the DAG on the GPU
(reference ``analytic/_propagator.py``).

``run()`` flattens the context into arrays -- integer bounds of every event, the precedence list, one PMF per
edge -- and makes ONE call into the library (``mcdp_analytic_run``): a kernel launch per topological level, a CTA per
event, convolution / maximum / bound handling block-cooperative in compensated (twice-working-precision) arithmetic.  The result comes back
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
    _flat: tuple | None = None  # the context as device-call arrays (filled by the first run)

    @property
    def underflow_rule(self) -> UnderflowRule:
        return self.context.underflow_rule

    @property
    def overflow_rule(self) -> OverflowRule:
        return self.context.overflow_rule

    def _event_bounds(self, earliest: Second, latest: Second) -> tuple[int, int]:
        """Integer bounds of one event (reference ``_propagator.py:150-156``); ``_flatten`` does all events at once."""
        upper = latest
        if self.context.max_delay is not None:
            upper = min(latest, earliest + self.context.max_delay)
        return int(np.round(earliest)), int(np.round(upper))

    def _flatten(self) -> tuple:
        """The context as the arrays ``mcdp_analytic_run`` takes: integer bounds and origins of every event, the
        precedence list in CSR form, one PMF per distinct activity object in first-use order.  Contexts are immutable,
        so this is done once per propagator."""
        if self._flat is not None:
            return self._flat
        ctx = self.context
        step = ctx.step
        earliest = np.array([ev.timestamp.earliest for ev in ctx.events], dtype=float)
        latest = np.array([ev.timestamp.latest for ev in ctx.events], dtype=float)
        if ctx.max_delay is not None:
            latest = np.minimum(latest, earliest + ctx.max_delay)
        lower, upper = np.round(earliest).astype(np.int64), np.round(latest).astype(np.int64)
        origin = np.round(np.round(earliest / step) * step).astype(np.int64)
        pmf_index: dict[int, int] = {}
        pmf_start: list[int] = []
        pmf_probs: list[np.ndarray] = []
        targets, offsets, srcs, pmfs = [], [0], [], []
        for target, preds in enumerate(self._predecessors_by_target):
            if not preds:
                continue
            targets.append(target)
            for src, _ in preds:
                pmf = ctx.activities[(src, target)][1].pmf
                slot = pmf_index.get(id(pmf))
                if slot is None:
                    slot = pmf_index[id(pmf)] = len(pmf_start)
                    pmf_start.append(int(round(float(pmf.values[0]))))
                    pmf_probs.append(np.ascontiguousarray(pmf.probabilities, np.float64))
                srcs.append(src)
                pmfs.append(slot)
            offsets.append(len(srcs))
        pmf_off = np.zeros(len(pmf_probs) + 1, np.int64)
        np.cumsum([p.size for p in pmf_probs], out=pmf_off[1:])
        flat = (lower, upper, origin, int(step), np.asarray(targets, np.int32), np.asarray(offsets, np.int64),
                np.asarray(srcs, np.int32), np.asarray(pmfs, np.int32), np.asarray(pmf_start, np.int64), pmf_off,
                np.concatenate(pmf_probs) if pmf_probs else np.zeros(0))
        object.__setattr__(self, "_flat", flat)
        return flat

    def run(self) -> tuple[SimulatedEvent, ...]:
        """One ``SimulatedEvent`` per event, in the order of ``context.events``."""
        step = self.context.step
        start, length, off, probs, under, over = _device.analytic_run(
            *self._flatten(), int(self.underflow_rule), int(self.overflow_rule))
        n = len(length)
        if n == 0:
            return ()
        # results packed back to back (`off` are the result slots, sized from the event bounds, usually longer than the
        # results) and the value grids of all events in one array; both are sliced per event below
        vend = np.cumsum(length, dtype=np.int64)
        vbeg = vend - length
        within = np.arange(int(vend[-1]), dtype=np.int64) - np.repeat(vbeg, length)
        if int(off[-1]) != int(vend[-1]):
            probs = probs[np.repeat(off[:-1], length) + within]
        values = np.repeat(start.astype(float), length) + float(step) * within.astype(float)
        # the checks of the reference's DiscretePMF constructor (_pmf.py:38-52) for all events at once: the grids are
        # ascending by construction, every event has at least one bin, what is left to check is the mass
        mass = np.add.reduceat(probs, vbeg)
        if np.any((mass > 1.0) & ~np.isclose(mass, 1.0)):
            raise ValueError("Probabilities must sum to <= 1.0")
        begs, ends, unders, overs = vbeg.tolist(), vend.tolist(), under.tolist(), over.tolist()  # Python scalars in the loop
        unchecked = DiscretePMF._unchecked
        out = []
        for i in range(n):
            v, e = begs[i], ends[i]
            out.append(SimulatedEvent(unchecked(values[v:e], probs[v:e], step), unders[i], overs[i]))
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
