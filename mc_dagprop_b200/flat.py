"""Flat (array) descriptions of a DAG and of a delay generator.

These are the host-side data formats either side of the hot path: what the reference keeps in
``DagContext`` (``_core.cpp:52-62``: ``vector<Event>``, ``unordered_map<(src,dst), Activity>``,
``precedence_list``) and in ``GenericDelayGenerator::dist_map_`` (``_core.cpp:146-159``), as plain
numpy arrays that cross the C ABI (``include/mcdp_b200.h``) without per-object conversion.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

KIND_CONSTANT, KIND_EXPONENTIAL, KIND_GAMMA, KIND_EMP_ABS, KIND_EMP_REL = 0, 1, 2, 3, 4


@dataclass
class FlatDag:
    """``earliest[E]``; activities as parallel arrays (``act_idx`` is ``Activity.idx``, gaps allowed);
    precedence as the reference's ``precedence_list`` flattened: entry ``i`` targets
    ``prec_target[i]`` with predecessors ``(pred_src[k], pred_act[k])`` for
    ``k in prec_off[i]:prec_off[i+1]`` -- in the caller's order, which decides ties."""

    earliest: np.ndarray
    act_idx: np.ndarray
    act_base: np.ndarray
    act_type: np.ndarray
    prec_target: np.ndarray
    prec_off: np.ndarray
    pred_src: np.ndarray
    pred_act: np.ndarray
    max_delay: float

    def __post_init__(self):
        self.earliest = np.ascontiguousarray(self.earliest, np.float64).reshape(-1)
        self.act_idx = np.ascontiguousarray(self.act_idx, np.int32).reshape(-1)
        self.act_base = np.ascontiguousarray(self.act_base, np.float64).reshape(-1)
        self.act_type = np.ascontiguousarray(self.act_type, np.int32).reshape(-1)
        self.prec_target = np.ascontiguousarray(self.prec_target, np.int32).reshape(-1)
        self.prec_off = np.ascontiguousarray(self.prec_off, np.int64).reshape(-1)
        if self.prec_off.size == 0:
            self.prec_off = np.zeros(1, np.int64)
        self.pred_src = np.ascontiguousarray(self.pred_src, np.int32).reshape(-1)
        self.pred_act = np.ascontiguousarray(self.pred_act, np.int32).reshape(-1)
        self.max_delay = float(self.max_delay)

    @property
    def n_events(self) -> int:
        return int(self.earliest.size)

    @property
    def n_activities(self) -> int:
        """``activity_count()`` of the reference: max idx + 1 (``_core.cpp:213-217``)."""
        return int(self.act_idx.max()) + 1 if self.act_idx.size else 0

    @property
    def n_preds(self) -> int:
        return int(self.pred_src.size)

    @classmethod
    def from_precedence_list(cls, earliest, activities, precedence_list, max_delay) -> "FlatDag":
        """``activities``: iterable of ``(idx, minimal_duration, activity_type)``;
        ``precedence_list``: ``[(target, [(src, act_idx), ...]), ...]`` as in the reference API."""
        acts = list(activities)
        tgt, off, src, act = [], [0], [], []
        for t, preds in precedence_list:
            tgt.append(t)
            for s, a in preds:
                src.append(s)
                act.append(a)
            off.append(len(src))
        return cls(
            earliest=np.asarray(earliest, np.float64),
            act_idx=np.asarray([a[0] for a in acts], np.int32),
            act_base=np.asarray([a[1] for a in acts], np.float64),
            act_type=np.asarray([a[2] for a in acts], np.int32),
            prec_target=np.asarray(tgt, np.int32),
            prec_off=np.asarray(off, np.int64),
            pred_src=np.asarray(src, np.int32),
            pred_act=np.asarray(act, np.int32),
            max_delay=max_delay,
        )


@dataclass
class FlatDists:
    """Parameter tables of a ``GenericDelayGenerator``; same ``add_*`` names and argument
    meaning as the reference binding (``_core.cpp:519-542``).  A later ``add_*`` for the same
    ``activity_type`` replaces the earlier one."""

    _entries: dict = field(default_factory=dict)

    def add_constant(self, activity_type: int, factor: float) -> None:
        self._entries[int(activity_type)] = (KIND_CONSTANT, float(factor), 0.0, 0.0, None, None)

    def add_exponential(self, activity_type: int, lambda_: float, max_scale: float) -> None:
        self._entries[int(activity_type)] = (KIND_EXPONENTIAL, float(lambda_), float(max_scale), 0.0, None, None)

    def add_gamma(self, activity_type: int, shape: float, scale: float, max_scale: float = float("inf")) -> None:
        self._entries[int(activity_type)] = (KIND_GAMMA, float(shape), float(scale), float(max_scale), None, None)

    def _add_table(self, kind, activity_type, values, weights):
        v = np.asarray(values, np.float64).reshape(-1).copy()
        w = np.asarray(weights, np.float64).reshape(-1).copy()
        if v.size != w.size:
            name = "EmpiricalAbsoluteDist: values" if kind == KIND_EMP_ABS else "EmpiricalRelativeDist: factors"
            raise RuntimeError(f"{name} and weights must have same length")
        self._entries[int(activity_type)] = (kind, 0.0, 0.0, 0.0, v, w)

    def add_empirical_absolute(self, activity_type: int, values, weights) -> None:
        self._add_table(KIND_EMP_ABS, activity_type, values, weights)

    def add_empirical_relative(self, activity_type: int, factors, weights) -> None:
        self._add_table(KIND_EMP_REL, activity_type, factors, weights)

    # flattened views -------------------------------------------------------
    def _flatten(self):
        types = list(self._entries.keys())
        kind = [self._entries[t][0] for t in types]
        p0 = [self._entries[t][1] for t in types]
        p1 = [self._entries[t][2] for t in types]
        p2 = [self._entries[t][3] for t in types]
        off, vals, wts = [0], [], []
        for t in types:
            v, w = self._entries[t][4], self._entries[t][5]
            if v is not None:
                vals.append(v)
                wts.append(w)
            off.append(off[-1] + (0 if v is None else v.size))
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.float64)  # noqa: E731
        return (np.asarray(types, np.int32), np.asarray(kind, np.int32), np.asarray(p0, np.float64),
                np.asarray(p1, np.float64), np.asarray(p2, np.float64), np.asarray(off, np.int64), cat(vals), cat(wts))

    dist_type = property(lambda self: self._flatten()[0])
    kind = property(lambda self: self._flatten()[1])
    p0 = property(lambda self: self._flatten()[2])
    p1 = property(lambda self: self._flatten()[3])
    p2 = property(lambda self: self._flatten()[4])
    tab_off = property(lambda self: self._flatten()[5])
    tab_values = property(lambda self: self._flatten()[6])
    tab_weights = property(lambda self: self._flatten()[7])
