"""ctypes seam of the analytic engine: ``mcdp_pmf_op`` / ``mcdp_analytic_run`` of ``libmcdp_b200.so``.

The arithmetic of the analytic propagator (convolution, maximum, bound handling) runs on the GPU
(``csrc/mcdp_analytic.cu``); like the Monte-Carlo path it has no CPU execution path -- without a CUDA device
these calls raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import capi

_DEVICE = 0  # CUDA ordinal the analytic engine runs on (set_device)


def set_device(device: int) -> None:
    """Choose the CUDA device of the analytic engine (default 0)."""
    global _DEVICE
    _DEVICE = int(device)


def get_device() -> int:
    return _DEVICE


class AnalyticDesc(C.Structure):
    _fields_ = [
        ("n_events", C.c_int32), ("lower", C.c_void_p), ("upper", C.c_void_p), ("origin", C.c_void_p), ("step", C.c_int64),
        ("n_prec_entries", C.c_int32), ("prec_target", C.c_void_p), ("prec_off", C.c_void_p), ("pred_src", C.c_void_p),
        ("pred_pmf", C.c_void_p), ("n_pmfs", C.c_int32), ("pmf_start", C.c_void_p), ("pmf_off", C.c_void_p),
        ("pmf_probs", C.c_void_p), ("underflow_rule", C.c_int32), ("overflow_rule", C.c_int32),
    ]


_bound = False


def _lib():
    global _bound
    L = capi.lib()
    if not _bound:
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.mcdp_analytic_out_capacity.argtypes = [C.POINTER(AnalyticDesc), vp]
        L.mcdp_analytic_out_capacity.restype = i64
        L.mcdp_analytic_run.argtypes = [C.POINTER(AnalyticDesc), i32, vp, vp, vp, vp, i64, vp, vp]
        L.mcdp_analytic_last_profile.argtypes = [C.POINTER(C.c_double)]
        L.mcdp_pmf_op.argtypes = [i32, i32, i64, i64, i32, vp, i64, i32, vp, i64, i64, i32, i32, vp, vp, vp, i64, vp, vp]
        _bound = True
    return L


def _raise(rc: int) -> None:
    msg = capi.lib().mcdp_last_error().decode()
    # what the reference reports as ValueError (validation, bound handling) stays ValueError; the rest is RuntimeError
    if rc == capi.MCDP_ERR_INVALID and "cycle" not in msg:
        raise ValueError(msg)
    raise RuntimeError(msg)


def last_profile() -> dict:
    """Phases of this thread's last ``analytic_run`` in milliseconds (``mcdp_analytic_last_profile``)."""
    out = (C.c_double * 5)()
    _lib().mcdp_analytic_last_profile(out)
    return {"host_prepare_ms": out[0], "alloc_upload_ms": out[1], "level_kernels_ms": out[2], "download_ms": out[3],
            "levels": int(out[4])}


def pmf_op(op: int, step: int, a_start: int, a_probs: np.ndarray, b_start: int = 0, b_probs: np.ndarray | None = None,
           bounds: tuple[int, int] = (0, 0), rules: tuple[int, int] = (1, 1)):
    """One PMF operation on the device: 0 convolve, 1 maximum, 2 clip.  Returns (start, probs, underflow, overflow)."""
    L = _lib()
    a = np.ascontiguousarray(a_probs, np.float64)
    b = np.ascontiguousarray(b_probs if b_probs is not None else np.zeros(1), np.float64)
    if op == 0:
        cap = a.size + b.size
    elif op == 1:
        s = max(int(step), 1)
        lo = min(a_start, b_start)
        hi = max(a_start + (a.size - 1) * s, b_start + (b.size - 1) * s)
        cap = (hi - lo) // s + 2
    else:
        cap = max(a.size, 1) + 1
    out = np.empty(cap, np.float64)
    o_start, o_len = C.c_int64(0), C.c_int32(0)
    under, over = C.c_double(0.0), C.c_double(0.0)
    rc = L.mcdp_pmf_op(op, _DEVICE, int(step), int(a_start), a.size, a.ctypes.data, int(b_start), b.size if op < 2 else 0,
                       b.ctypes.data if op < 2 else None, int(bounds[0]), int(bounds[1]), int(rules[0]), int(rules[1]),
                       C.addressof(o_start), C.addressof(o_len), out.ctypes.data, cap, C.addressof(under), C.addressof(over))
    if rc:
        _raise(rc)
    return int(o_start.value), out[: o_len.value].copy(), float(under.value), float(over.value)


def analytic_run(lower, upper, origin, step, prec_target, prec_off, pred_src, pred_pmf, pmf_start, pmf_off, pmf_probs,
                 underflow_rule: int, overflow_rule: int):
    """The whole DAG in one call.  Returns (start[E], len[E], off[E+1], probs, underflow[E], overflow[E])."""
    L = _lib()
    i32, i64, f64 = np.int32, np.int64, np.float64
    keep = [np.ascontiguousarray(x, dt) for x, dt in (
        (lower, i64), (upper, i64), (origin, i64), (prec_target, i32), (prec_off, i64), (pred_src, i32), (pred_pmf, i32),
        (pmf_start, i64), (pmf_off, i64), (pmf_probs, f64))]
    for k in (4, 8):
        if keep[k].size == 0:
            keep[k] = np.zeros(1, i64)
    E = keep[0].size
    d = AnalyticDesc(E, keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data, int(step), keep[3].size,
                     keep[3].ctypes.data, keep[4].ctypes.data, keep[5].ctypes.data, keep[6].ctypes.data, keep[8].size - 1,
                     keep[7].ctypes.data, keep[8].ctypes.data, keep[9].ctypes.data, int(underflow_rule), int(overflow_rule))
    cap = int(L.mcdp_analytic_out_capacity(C.byref(d), None))
    if cap < 0:
        raise ValueError("step_size must be positive")
    o_start, o_len, o_off = np.zeros(E, i64), np.zeros(E, i32), np.zeros(E + 1, i64)
    probs = np.zeros(max(cap, 1), f64)
    under, over = np.zeros(E, f64), np.zeros(E, f64)
    rc = L.mcdp_analytic_run(C.byref(d), _DEVICE, o_start.ctypes.data, o_len.ctypes.data, o_off.ctypes.data, probs.ctypes.data, cap,
                             under.ctypes.data, over.ctypes.data)
    if rc:
        _raise(rc)
    return o_start, o_len, o_off, probs, under, over
