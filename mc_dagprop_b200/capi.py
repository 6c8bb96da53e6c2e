"""ctypes binding of ``libmcdp_b200.so`` (C ABI in ``include/mcdp_b200.h``).

This is the thin seam the parity tests and ``bench.py`` call through.  There is no CPU
execution path behind it: if the shared library is missing the import of :func:`lib` raises,
and without a CUDA device every run call raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmcdp_b200.so")

MCDP_OK, MCDP_ERR_INVALID, MCDP_ERR_CUDA, MCDP_ERR_ARG = 0, 1, 2, 3
DEVICE_NONE = -1
OPT_STREAM_KEY, OPT_WARPS_PER_GROUP, OPT_GROUPS_PER_CTA, OPT_HOST_CHUNK, OPT_RNG_STREAM, OPT_SAMPLES_PER_LANE = 0, 1, 2, 3, 4, 5
OPT_CLUSTER_SIZE = 6
OPT_SMALL_CALL_MAX = 7
RNG_PHILOX, RNG_REFERENCE = 0, 1
MAX_THRESHOLDS = 4
CHUNK_UNITS = 16
KIND_NONE, KIND_EVENT, KIND_END, NO_ROW, NO_ACT = 5, 6, 7, 0xFFFFFFFF, 0xFFFFFFFF
#: one 32-byte unit of the chunk stream, viewed as a precedence entry (csrc/mcdp_records.h: PredRec); an event header
#: (HeaderUnit) reads a = row, b = event id, x = earliest, meta, c = remaining, nxt = first_src_row, d = forward flag
UNIT_DTYPE = np.dtype([("a", "<u4"), ("b", "<u4"), ("x", "<f8"), ("meta", "<u4"), ("c", "<u4"), ("nxt", "<u4"), ("d", "<u4")])

#: every symbol include/mcdp_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = (
    "mcdp_last_error", "mcdp_abi_version", "mcdp_device_count", "mcdp_plan_create", "mcdp_plan_destroy",
    "mcdp_plan_set_option", "mcdp_plan_node_count", "mcdp_plan_activity_count", "mcdp_plan_pred_count",
    "mcdp_plan_level_count", "mcdp_plan_slot_count", "mcdp_plan_device", "mcdp_plan_get_order", "mcdp_plan_get_cumulative",
    "mcdp_run_full_device", "mcdp_run_injected_device", "mcdp_run_reduced_device", "mcdp_transpose_f64_device",
    "mcdp_transpose_i32_device", "mcdp_run_many_host", "mcdp_run_injected_host", "mcdp_run_reduced_host",
    "mcdp_plan_get_chunks", "mcdp_plan_launch_shape", "mcdp_plan_reduced_chunk", "mcdp_run_attribution_device", "mcdp_run_attribution_host", "mcdp_host_alloc", "mcdp_host_free",
    "mcdp_planset_create", "mcdp_planset_destroy", "mcdp_planset_size", "mcdp_planset_plan", "mcdp_planset_set_option",
    "mcdp_run_many_host_multi", "mcdp_run_injected_host_multi", "mcdp_run_reduced_host_multi", "mcdp_run_attribution_host_multi",
    "mcdp_analytic_out_capacity", "mcdp_analytic_run", "mcdp_analytic_last_profile", "mcdp_pmf_op",
)


class GraphDesc(C.Structure):
    _fields_ = [
        ("n_events", C.c_int32), ("earliest", C.c_void_p),
        ("n_act_entries", C.c_int32), ("act_idx", C.c_void_p), ("act_base", C.c_void_p), ("act_type", C.c_void_p),
        ("n_prec_entries", C.c_int32), ("prec_target", C.c_void_p), ("prec_off", C.c_void_p),
        ("pred_src", C.c_void_p), ("pred_act", C.c_void_p), ("max_delay", C.c_double),
    ]


class DistsDesc(C.Structure):
    _fields_ = [
        ("n_dists", C.c_int32), ("dist_type", C.c_void_p), ("kind", C.c_void_p), ("p0", C.c_void_p),
        ("p1", C.c_void_p), ("p2", C.c_void_p), ("tab_off", C.c_void_p), ("tab_values", C.c_void_p),
        ("tab_weights", C.c_void_p),
    ]


class StatsDesc(C.Structure):
    _fields_ = [
        ("n_thresholds", C.c_int32), ("thresholds", C.c_double * MAX_THRESHOLDS), ("n_bins", C.c_int32),
        ("hist_lo", C.c_double), ("hist_hi", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the native library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m mc_dagprop_b200.build` "
                "(mc_dagprop_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.mcdp_last_error.restype = C.c_char_p
        L.mcdp_plan_create.argtypes = [C.POINTER(GraphDesc), C.POINTER(DistsDesc), i32, C.POINTER(vp)]
        L.mcdp_plan_destroy.argtypes = [vp]
        L.mcdp_plan_destroy.restype = None
        L.mcdp_plan_set_option.argtypes = [vp, i32, i64]
        for name in ("node_count", "activity_count", "level_count", "slot_count", "device"):
            getattr(L, f"mcdp_plan_{name}").argtypes = [vp]
        L.mcdp_plan_pred_count.argtypes = [vp]
        L.mcdp_plan_pred_count.restype = i64
        L.mcdp_plan_get_order.argtypes = [vp, vp, vp]
        L.mcdp_plan_get_cumulative.argtypes = [vp, i32, vp, i64]
        L.mcdp_plan_get_cumulative.restype = i64
        L.mcdp_plan_get_chunks.argtypes = [vp, i32, i32, vp, i64, vp]
        L.mcdp_plan_get_chunks.restype = i64
        L.mcdp_plan_launch_shape.argtypes = [vp, i64, i32, i32, vp]
        L.mcdp_plan_reduced_chunk.argtypes = [vp, i64, i32, i32]
        L.mcdp_plan_reduced_chunk.restype = i64
        L.mcdp_run_full_device.argtypes = [vp, vp, i32, i64, vp, vp, vp, i64, vp]
        L.mcdp_run_injected_device.argtypes = [vp, vp, i64, vp, vp, i64, vp]
        L.mcdp_run_reduced_device.argtypes = [vp, vp, i32, i64, C.POINTER(StatsDesc), vp, vp, vp, vp, vp]
        L.mcdp_transpose_f64_device.argtypes = [vp, i64, i64, i64, vp, vp]
        L.mcdp_transpose_i32_device.argtypes = [vp, i64, i64, i64, vp, vp]
        L.mcdp_run_many_host.argtypes = [vp, vp, i64, vp, vp, vp]
        L.mcdp_run_injected_host.argtypes = [vp, vp, i64, vp, vp]
        L.mcdp_run_reduced_host.argtypes = [vp, vp, i64, C.POINTER(StatsDesc), vp, vp, vp, vp]
        L.mcdp_run_attribution_device.argtypes = [vp, vp, i32, i64, C.POINTER(StatsDesc), vp, vp, vp, vp, vp, vp, vp]
        L.mcdp_run_attribution_host.argtypes = [vp, vp, i64, C.POINTER(StatsDesc), vp, vp, vp, vp, vp, vp]
        L.mcdp_planset_create.argtypes = [C.POINTER(GraphDesc), C.POINTER(DistsDesc), vp, i32, C.POINTER(vp)]
        L.mcdp_planset_destroy.argtypes = [vp]
        L.mcdp_planset_destroy.restype = None
        L.mcdp_planset_size.argtypes = [vp]
        L.mcdp_planset_plan.argtypes = [vp, i32]
        L.mcdp_planset_plan.restype = vp
        L.mcdp_planset_set_option.argtypes = [vp, i32, i64]
        L.mcdp_run_many_host_multi.argtypes = [vp, vp, i64, vp, vp, vp]
        L.mcdp_run_injected_host_multi.argtypes = [vp, vp, i64, vp, vp]
        L.mcdp_run_reduced_host_multi.argtypes = [vp, vp, i64, C.POINTER(StatsDesc), vp, vp, vp, vp]
        L.mcdp_run_attribution_host_multi.argtypes = [vp, vp, i64, C.POINTER(StatsDesc), vp, vp, vp, vp, vp, vp]
        L.mcdp_host_alloc.argtypes = [C.c_size_t]
        L.mcdp_host_alloc.restype = vp
        L.mcdp_host_free.argtypes = [vp]
        L.mcdp_host_free.restype = None
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != MCDP_OK:
        raise RuntimeError(lib().mcdp_last_error().decode())


def device_count() -> int:
    return int(lib().mcdp_device_count())


def _ptr(a) -> int | None:
    """numpy array, torch tensor, int address or None -> address."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)!r}")


def _np(x, dt):
    return np.ascontiguousarray(np.asarray(x, dtype=dt).reshape(-1))


@dataclass
class Stats:
    """Per-event statistics of ``realized - earliest`` accumulated by the reduced mode."""

    n: int
    sum: np.ndarray
    sumsq: np.ndarray
    late: np.ndarray          # [n_thresholds, E]
    hist: np.ndarray          # [E, n_bins]
    thresholds: tuple = ()
    hist_range: tuple = (0.0, 0.0)

    @property
    def mean(self):
        return self.sum / max(self.n, 1)

    @property
    def var(self):
        m = self.mean
        return np.maximum(self.sumsq / max(self.n, 1) - m * m, 0.0)

    def quantile(self, q: float) -> np.ndarray:
        """Per-event delay quantile from the histogram (upper bin edge)."""
        lo, hi = self.hist_range
        nb = self.hist.shape[1]
        cdf = np.cumsum(self.hist, axis=1)
        k = (cdf >= q * self.n).argmax(axis=1)
        return lo + (k + 1) * (hi - lo) / nb


def make_stats_desc(thresholds=(), n_bins=0, hist_range=(0.0, 1.0)) -> StatsDesc:
    d = StatsDesc()
    th = tuple(float(t) for t in thresholds)
    if len(th) > MAX_THRESHOLDS:
        raise ValueError(f"at most {MAX_THRESHOLDS} thresholds")
    d.n_thresholds = len(th)
    for i, t in enumerate(th):
        d.thresholds[i] = t
    d.n_bins = int(n_bins)
    d.hist_lo, d.hist_hi = float(hist_range[0]), float(hist_range[1])
    return d


def _descs(dag, dists):
    """Flat descriptions -> (GraphDesc, DistsDesc, arrays that must stay alive during the call)."""
    keep = [
        _np(dag.earliest, np.float64), _np(dag.act_idx, np.int32), _np(dag.act_base, np.float64),
        _np(dag.act_type, np.int32), _np(dag.prec_target, np.int32), _np(dag.prec_off, np.int64),
        _np(dag.pred_src, np.int32), _np(dag.pred_act, np.int32),
        _np(dists.dist_type, np.int32), _np(dists.kind, np.int32), _np(dists.p0, np.float64),
        _np(dists.p1, np.float64), _np(dists.p2, np.float64), _np(dists.tab_off, np.int64),
        _np(dists.tab_values, np.float64), _np(dists.tab_weights, np.float64),
    ]
    if keep[5].size == 0:
        keep[5] = np.zeros(1, np.int64)
    if keep[13].size == 0:
        keep[13] = np.zeros(1, np.int64)
    g = GraphDesc(keep[0].size, keep[0].ctypes.data, keep[1].size, keep[1].ctypes.data, keep[2].ctypes.data,
                  keep[3].ctypes.data, keep[4].size, keep[4].ctypes.data, keep[5].ctypes.data,
                  keep[6].ctypes.data, keep[7].ctypes.data, float(dag.max_delay))
    d = DistsDesc(keep[8].size, keep[8].ctypes.data, keep[9].ctypes.data, keep[10].ctypes.data,
                  keep[11].ctypes.data, keep[12].ctypes.data, keep[13].ctypes.data, keep[14].ctypes.data,
                  keep[15].ctypes.data)
    return g, d, keep


class Plan:
    """A compiled, device-resident DAG + generator (``mcdp_plan``): the counterpart of a constructed
    reference ``Simulator`` (``_core.cpp:193-307``)."""

    def __init__(self, dag, dists, device: int = 0):
        L = lib()
        g, d, _keep = _descs(dag, dists)
        h = C.c_void_p()
        self._h = None
        _check(L.mcdp_plan_create(C.byref(g), C.byref(d), int(device), C.byref(h)))
        self._h = h
        self.E = int(L.mcdp_plan_node_count(h))
        self.A = int(L.mcdp_plan_activity_count(h))
        self.P = int(L.mcdp_plan_pred_count(h))
        self.n_levels = int(L.mcdp_plan_level_count(h))
        self.n_slots = int(L.mcdp_plan_slot_count(h))
        self.device = int(L.mcdp_plan_device(h))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:  # (_lib is None again at interpreter shutdown)
            _lib.mcdp_plan_destroy(self._h)
            self._h = None

    __del__ = close

    # -- introspection ------------------------------------------------------
    def order(self):
        o, lv = np.empty(max(self.E, 1), np.int32), np.empty(max(self.E, 1), np.int32)
        _check(lib().mcdp_plan_get_order(self._h, o.ctypes.data, lv.ctypes.data))
        return o[: self.E], lv[: self.E]

    def chunks(self, rows: int = 0, dense: bool = False):
        """The chunk stream the sweep kernel walks: (units[n_chunks, 16] structured array, chunk_level_begin)."""
        n_levels = 1 if dense else self.n_levels
        clb = np.zeros(n_levels + 1, np.int32)
        n = int(lib().mcdp_plan_get_chunks(self._h, int(rows), int(bool(dense)), None, 0, None))
        raw = np.zeros(max(n, 0) * CHUNK_UNITS * 32, np.uint8)
        lib().mcdp_plan_get_chunks(self._h, int(rows), int(bool(dense)), raw.ctypes.data if raw.size else None, raw.size,
                                   clb.ctypes.data)
        return raw.view(UNIT_DTYPE).reshape(n, CHUNK_UNITS), clb

    def cumulative(self, activity_type: int, cap: int = 1 << 20):
        out = np.empty(cap, np.float64)
        n = lib().mcdp_plan_get_cumulative(self._h, int(activity_type), out.ctypes.data, cap)
        return None if n < 0 else out[:n].copy()

    def launch_shape(self, n: int, reduced: bool = False, n_bins: int = 0) -> dict:
        """The launch a call over n samples would take (which kernel, CTA shape); works on host-only plans."""
        out = np.zeros(8, np.int64)
        _check(lib().mcdp_plan_launch_shape(self._h, int(n), int(bool(reduced)), int(n_bins), out.ctypes.data))
        keys = ("samples_per_lane", "warps_per_group", "groups_per_cta", "threads", "grid", "smem_bytes", "cluster", "smem_tables")
        return dict(zip(keys, out.tolist()))

    def reduced_chunk(self, n: int, n_bins: int = 0, attribution: bool = False) -> int:
        """Samples per launch a reduced call over n samples would use (scratch budget, whole waves)."""
        return int(lib().mcdp_plan_reduced_chunk(self._h, int(n), int(n_bins), int(bool(attribution))))

    def set_option(self, option: int, value: int) -> None:
        _check(lib().mcdp_plan_set_option(self._h, option, int(value)))

    # -- host-buffer calls (sample-major results, like n SimResult objects) ---
    def run_many_host(self, seeds, realized=True, durations=True, cause=True, out=None):
        seeds = _np(seeds, np.int32)
        n = seeds.size
        if out is None:
            r = np.empty((n, self.E), np.float64) if realized else None
            d = np.empty((n, self.A), np.float64) if durations else None
            c = np.empty((n, self.E), np.int32) if cause else None
        else:
            r, d, c = out
        _check(lib().mcdp_run_many_host(self._h, seeds.ctypes.data, n, _ptr(r), _ptr(d), _ptr(c)))
        return r, d, c

    def run_injected_host(self, durations):
        durations = np.ascontiguousarray(durations, np.float64)
        n = durations.shape[0]
        assert durations.size == n * self.A
        r, c = np.empty((n, self.E), np.float64), np.empty((n, self.E), np.int32)
        _check(lib().mcdp_run_injected_host(self._h, durations.ctypes.data if durations.size else None, n,
                                            r.ctypes.data, c.ctypes.data))
        return r, c

    def run_reduced_host(self, seeds, thresholds=(), n_bins=0, hist_range=(0.0, 1.0)) -> Stats:
        seeds = _np(seeds, np.int32)
        desc = make_stats_desc(thresholds, n_bins, hist_range)
        s, q = np.zeros(self.E, np.float64), np.zeros(self.E, np.float64)
        late = np.zeros((len(thresholds), self.E), np.uint64)
        hist = np.zeros((self.E, n_bins), np.uint32)
        _check(lib().mcdp_run_reduced_host(self._h, seeds.ctypes.data, seeds.size, C.byref(desc), s.ctypes.data,
                                           q.ctypes.data, late.ctypes.data if late.size else None,
                                           hist.ctypes.data if hist.size else None))
        return Stats(seeds.size, s, q, late, hist, tuple(thresholds), tuple(hist_range))

    def run_attribution_host(self, seeds, thresholds=(), n_bins=0, hist_range=(0.0, 1.0)):
        """Statistics + delay-cause attribution: returns (Stats, cause_act[A] u64, cause_none[E] u64)."""
        seeds = _np(seeds, np.int32)
        desc = make_stats_desc(thresholds, n_bins, hist_range)
        s, q = np.zeros(self.E, np.float64), np.zeros(self.E, np.float64)
        late = np.zeros((len(thresholds), self.E), np.uint64)
        hist = np.zeros((self.E, n_bins), np.uint32)
        cause_act, cause_none = np.zeros(self.A, np.uint64), np.zeros(self.E, np.uint64)
        _check(lib().mcdp_run_attribution_host(self._h, seeds.ctypes.data, seeds.size, C.byref(desc), s.ctypes.data,
                                               q.ctypes.data, late.ctypes.data if late.size else None,
                                               hist.ctypes.data if hist.size else None, cause_act.ctypes.data,
                                               cause_none.ctypes.data))
        return Stats(seeds.size, s, q, late, hist, tuple(thresholds), tuple(hist_range)), cause_act, cause_none

    # -- device-buffer calls (event-major [rows][ld]; torch tensors, addresses or None) --
    def run_full_device(self, n, realized, durations, cause, ld, seeds=None, seed0=0, stream=None):
        _check(lib().mcdp_run_full_device(self._h, _ptr(seeds), int(seed0), int(n), _ptr(realized), _ptr(durations),
                                          _ptr(cause), int(ld), _ptr(stream)))

    def run_injected_device(self, n, durations, realized, cause, ld, stream=None):
        _check(lib().mcdp_run_injected_device(self._h, _ptr(durations), int(n), _ptr(realized), _ptr(cause), int(ld),
                                              _ptr(stream)))

    def run_reduced_device(self, n, desc: StatsDesc, sum=None, sumsq=None, late=None, hist=None, seeds=None, seed0=0,
                           stream=None):
        _check(lib().mcdp_run_reduced_device(self._h, _ptr(seeds), int(seed0), int(n), C.byref(desc), _ptr(sum),
                                             _ptr(sumsq), _ptr(late), _ptr(hist), _ptr(stream)))


class PlanSet:
    """One compiled plan per device (``mcdp_planset``): every call shards its seeds into contiguous blocks, one
    per device, and returns what a single device returns.  ``devices`` may repeat an ordinal."""

    def __init__(self, dag, dists, devices):
        L = lib()
        g, d, _keep = _descs(dag, dists)
        self.devices = [int(x) for x in devices]
        dev = np.asarray(self.devices, np.int32)
        h = C.c_void_p()
        self._h = None
        _check(L.mcdp_planset_create(C.byref(g), C.byref(d), dev.ctypes.data, dev.size, C.byref(h)))
        self._h = h
        p0 = L.mcdp_planset_plan(h, 0)
        self.E = int(L.mcdp_plan_node_count(p0))
        self.A = int(L.mcdp_plan_activity_count(p0))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.mcdp_planset_destroy(self._h)
            self._h = None

    __del__ = close

    def __len__(self):
        return int(lib().mcdp_planset_size(self._h))

    def set_option(self, option: int, value: int) -> None:
        _check(lib().mcdp_planset_set_option(self._h, option, int(value)))

    def run_many_host(self, seeds, realized=True, durations=True, cause=True, out=None):
        seeds = _np(seeds, np.int32)
        n = seeds.size
        if out is None:
            r = np.empty((n, self.E), np.float64) if realized else None
            d = np.empty((n, self.A), np.float64) if durations else None
            c = np.empty((n, self.E), np.int32) if cause else None
        else:
            r, d, c = out
        _check(lib().mcdp_run_many_host_multi(self._h, seeds.ctypes.data, n, _ptr(r), _ptr(d), _ptr(c)))
        return r, d, c

    def run_injected_host(self, durations):
        durations = np.ascontiguousarray(durations, np.float64)
        n = durations.shape[0]
        assert durations.size == n * self.A
        r, c = np.empty((n, self.E), np.float64), np.empty((n, self.E), np.int32)
        _check(lib().mcdp_run_injected_host_multi(self._h, durations.ctypes.data if durations.size else None, n,
                                                  r.ctypes.data, c.ctypes.data))
        return r, c

    def run_attribution_host(self, seeds, thresholds=(), n_bins=0, hist_range=(0.0, 1.0), cause_counts=True):
        seeds = _np(seeds, np.int32)
        desc = make_stats_desc(thresholds, n_bins, hist_range)
        s, q = np.zeros(self.E, np.float64), np.zeros(self.E, np.float64)
        late = np.zeros((len(thresholds), self.E), np.uint64)
        hist = np.zeros((self.E, n_bins), np.uint32)
        cause_act = np.zeros(self.A, np.uint64) if cause_counts else None
        cause_none = np.zeros(self.E, np.uint64) if cause_counts else None
        _check(lib().mcdp_run_attribution_host_multi(self._h, seeds.ctypes.data, seeds.size, C.byref(desc), s.ctypes.data,
                                                     q.ctypes.data, late.ctypes.data if late.size else None,
                                                     hist.ctypes.data if hist.size else None, _ptr(cause_act), _ptr(cause_none)))
        return Stats(seeds.size, s, q, late, hist, tuple(thresholds), tuple(hist_range)), cause_act, cause_none

    def run_reduced_host(self, seeds, thresholds=(), n_bins=0, hist_range=(0.0, 1.0)) -> Stats:
        return self.run_attribution_host(seeds, thresholds, n_bins, hist_range, cause_counts=False)[0]
