"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-ends for the two CPU checkers:

* :class:`OracleSim`  -- ``oracle/_build/libmcdp_oracle.so``: plain-C restatement of the reference
  hot path (``mcdp_oracle.c``; every function cites the reference file:line).  Parity PINNED by
  ``tests/test_oracle_pinned.py`` against the reference's golden vectors and ``oracle/_ref``.
* :class:`RefSim`     -- ``oracle/_ref/libmcdp_ref.so``: the UNMODIFIED reference engine
  (``/root/reference/src/mc_dagprop/monte_carlo/_core.cpp``) behind a flat-array C ABI
  (``ref_driver.cpp``).  Built in the container by ``oracle/Makefile``; travels to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this
package.  ``mc_dagprop_b200`` never does: the product path has no CPU fallback.

Both classes take a *flat DAG description*: any object with the attributes of
``mc_dagprop_b200.flat.FlatDag`` (``earliest, act_idx, act_base, act_type, prec_target, prec_off,
pred_src, pred_act, max_delay``) and one with those of ``FlatDists`` (``dist_type, kind, p0, p1, p2,
tab_off, tab_values, tab_weights``).
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import subprocess
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "_build", "libmcdp_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libmcdp_ref.so")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

_CREATE_ARGS = [
    C.c_int32, _f64p,  # events
    C.c_int32, _i32p, _f64p, _i32p,  # activities
    C.c_int32, _i32p, _i64p, _i32p, _i32p, C.c_double,  # precedence, max_delay
    C.c_int32, _i32p, _i32p, _f64p, _f64p, _f64p, _i64p, _f64p, _f64p,  # dists
]


def build(ref: bool | None = None) -> None:
    """Run ``make`` for the C restatement (always) and the reference builds (when the reference
    source tree is present, i.e. in the build container)."""
    targets = ["oracle"]
    if ref is None:
        ref = os.path.isdir("/root/reference/src/mc_dagprop/monte_carlo")
    if ref:
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", _HERE, *targets], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def _a(x, dt):
    return np.ascontiguousarray(np.asarray(x, dtype=dt).reshape(-1))


def _pad(x):
    return x if x.size else np.zeros(1, x.dtype)


def _flat_args(dag, dists):
    earliest = _a(dag.earliest, np.float64)
    act_idx, act_base, act_type = _a(dag.act_idx, np.int32), _a(dag.act_base, np.float64), _a(dag.act_type, np.int32)
    prec_target, prec_off = _a(dag.prec_target, np.int32), _a(dag.prec_off, np.int64)
    pred_src, pred_act = _a(dag.pred_src, np.int32), _a(dag.pred_act, np.int32)
    dist_type, kind = _a(dists.dist_type, np.int32), _a(dists.kind, np.int32)
    p0, p1, p2 = _a(dists.p0, np.float64), _a(dists.p1, np.float64), _a(dists.p2, np.float64)
    tab_off = _a(dists.tab_off, np.int64)
    tab_values, tab_weights = _a(dists.tab_values, np.float64), _a(dists.tab_weights, np.float64)
    if prec_off.size == 0:
        prec_off = np.zeros(1, np.int64)
    if tab_off.size == 0:
        tab_off = np.zeros(1, np.int64)
    return [
        earliest.size, _pad(earliest),
        act_idx.size, _pad(act_idx), _pad(act_base), _pad(act_type),
        prec_target.size, _pad(prec_target), prec_off, _pad(pred_src), _pad(pred_act), float(dag.max_delay),
        dist_type.size, _pad(dist_type), _pad(kind), _pad(p0), _pad(p1), _pad(p2), tab_off,
        _pad(tab_values), _pad(tab_weights),
    ]


class _SimBase:
    E: int
    A: int

    def _alloc(self, n, realized=True, durations=True, cause=True):
        r = np.empty((n, self.E), np.float64) if realized else None
        d = np.empty((n, self.A), np.float64) if durations else None
        c = np.empty((n, self.E), np.int32) if cause else None
        return r, d, c

    @staticmethod
    def _p(arr, ct):
        return None if arr is None else arr.ctypes.data_as(C.POINTER(ct))


_oracle_lib = None


def _load_oracle():
    global _oracle_lib
    if _oracle_lib is None:
        if not os.path.exists(ORACLE_LIB):
            build(ref=False)
        lib = C.CDLL(ORACLE_LIB)
        lib.mcdp_or_sim_create.restype = C.c_void_p
        lib.mcdp_or_sim_create.argtypes = _CREATE_ARGS + [C.c_char_p, C.c_size_t]
        lib.mcdp_or_sim_destroy.argtypes = [C.c_void_p]
        lib.mcdp_or_sim_node_count.argtypes = [C.c_void_p]
        lib.mcdp_or_sim_activity_count.argtypes = [C.c_void_p]
        lib.mcdp_or_sim_get_order.argtypes = [C.c_void_p, _i32p]
        lib.mcdp_or_sim_pred_count.argtypes = [C.c_void_p]
        lib.mcdp_or_sim_pred_count.restype = C.c_int64
        lib.mcdp_or_sim_get_csr.argtypes = [C.c_void_p, _i64p, _i32p, _i32p]
        lib.mcdp_or_sim_get_cp.argtypes = [C.c_void_p, C.c_int32, _f64p, C.c_int64]
        lib.mcdp_or_sim_get_cp.restype = C.c_int64
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        lib.mcdp_or_sim_run_many.argtypes = [C.c_void_p, _i32p, C.c_int64, dp, dp, ip]
        lib.mcdp_or_sim_run_injected.argtypes = [C.c_void_p, _f64p, C.c_int64, dp, ip]
        lib.mcdp_or_sim_run_many_spec.argtypes = [C.c_void_p, _i32p, C.c_int64, C.c_uint32, dp, dp, ip]
        lib.mcdp_or_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.mcdp_or_splitmix64_next.argtypes = [C.POINTER(C.c_uint64)]
        lib.mcdp_or_splitmix64_next.restype = C.c_uint64
        lib.mcdp_or_xoshiro_seed.argtypes = [C.c_void_p, C.c_uint64]
        lib.mcdp_or_xoshiro_next.argtypes = [C.c_void_p]
        lib.mcdp_or_xoshiro_next.restype = C.c_uint64
        lib.mcdp_or_canonical.argtypes = [C.c_void_p]
        lib.mcdp_or_canonical.restype = C.c_double
        _oracle_lib = lib
    return _oracle_lib


class OracleSim(_SimBase):
    """C restatement of ``Simulator`` (reference ``_core.cpp:162-362``)."""

    def __init__(self, dag, dists):
        self._lib = _load_oracle()
        err = C.create_string_buffer(256)
        self._h = self._lib.mcdp_or_sim_create(*_flat_args(dag, dists), err, 256)
        if not self._h:
            raise RuntimeError(err.value.decode())
        self.E = self._lib.mcdp_or_sim_node_count(self._h)
        self.A = self._lib.mcdp_or_sim_activity_count(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.mcdp_or_sim_destroy(self._h)
            self._h = None

    def order(self):
        out = np.empty(max(self.E, 1), np.int32)
        self._lib.mcdp_or_sim_get_order(self._h, out)
        return out[: self.E]

    def csr(self):
        P = self._lib.mcdp_or_sim_pred_count(self._h)
        off, src, act = np.empty(self.E + 1, np.int64), np.empty(max(P, 1), np.int32), np.empty(max(P, 1), np.int32)
        self._lib.mcdp_or_sim_get_csr(self._h, off, src, act)
        return off, src[:P], act[:P]

    def cumulative(self, dist_type, cap=1 << 20):
        out = np.empty(cap, np.float64)
        n = self._lib.mcdp_or_sim_get_cp(self._h, dist_type, out, cap)
        return out[: max(n, 0)].copy()

    def run_many(self, seeds, realized=True, durations=True, cause=True):
        """Reference stream (Xoshiro256++, activity-index order); sample-major outputs."""
        seeds = _a(seeds, np.int32)
        r, d, c = self._alloc(seeds.size, realized, durations, cause)
        self._lib.mcdp_or_sim_run_many(self._h, _pad(seeds), seeds.size, self._p(r, C.c_double),
                                       self._p(d, C.c_double), self._p(c, C.c_int32))
        return r, d, c

    def run_injected(self, durations):
        durations = np.ascontiguousarray(durations, np.float64).reshape(-1, max(self.A, 0)) if self.A else \
            np.zeros((np.asarray(durations).shape[0], 0))
        n = durations.shape[0]
        r, _, c = self._alloc(n, True, False, True)
        self._lib.mcdp_or_sim_run_injected(self._h, _pad(durations.reshape(-1)), n, self._p(r, C.c_double),
                                           self._p(c, C.c_int32))
        return r, c

    def run_many_spec(self, seeds, stream_key=0, realized=True, durations=True, cause=True):
        """Device generator contract mcdp-philox-v2 (not reference behaviour) + reference propagation."""
        seeds = _a(seeds, np.int32)
        r, d, c = self._alloc(seeds.size, realized, durations, cause)
        self._lib.mcdp_or_sim_run_many_spec(self._h, _pad(seeds), seeds.size, stream_key, self._p(r, C.c_double),
                                            self._p(d, C.c_double), self._p(c, C.c_int32))
        return r, d, c


def philox4x32_10(ctr, key):
    lib = _load_oracle()
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib.mcdp_or_philox4x32_10(c, k, o)
    return [int(x) for x in o]


class Xoshiro:
    """Reference RNG restatement, exposed for the known-answer tests."""

    def __init__(self, seed):
        self._lib = _load_oracle()
        self._s = (C.c_uint64 * 4)()
        self._lib.mcdp_or_xoshiro_seed(self._s, int(seed) & 0xFFFFFFFFFFFFFFFF)

    def next(self):
        return int(self._lib.mcdp_or_xoshiro_next(self._s))

    def canonical(self):
        return float(self._lib.mcdp_or_canonical(self._s))


_ref_lib = None


def _load_ref():
    global _ref_lib
    if _ref_lib is None:
        if not os.path.exists(REF_LIB):
            raise FileNotFoundError(f"{REF_LIB} missing: run `make -C oracle ref` in the build container")
        # PyDLL: the reference TU pulls in pybind11, whose symbols resolve against the interpreter.
        lib = C.CDLL(REF_LIB)
        lib.mcdp_ref_create.restype = C.c_void_p
        lib.mcdp_ref_create.argtypes = _CREATE_ARGS
        lib.mcdp_ref_destroy.argtypes = [C.c_void_p]
        lib.mcdp_ref_last_error.restype = C.c_char_p
        lib.mcdp_ref_node_count.argtypes = [C.c_void_p]
        lib.mcdp_ref_activity_count.argtypes = [C.c_void_p]
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        lib.mcdp_ref_run_many.argtypes = [C.c_void_p, _i32p, C.c_int64, dp, dp, ip]
        _ref_lib = lib
    return _ref_lib


class RefSim(_SimBase):
    """The unmodified reference ``Simulator`` (``oracle/_ref/libmcdp_ref.so``)."""

    def __init__(self, dag, dists):
        self._lib = _load_ref()
        self._h = self._lib.mcdp_ref_create(*_flat_args(dag, dists))
        if not self._h:
            raise RuntimeError(self._lib.mcdp_ref_last_error().decode())
        self.E = self._lib.mcdp_ref_node_count(self._h)
        self.A = self._lib.mcdp_ref_activity_count(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.mcdp_ref_destroy(self._h)
            self._h = None

    def run_many(self, seeds, realized=True, durations=True, cause=True):
        seeds = _a(seeds, np.int32)
        r, d, c = self._alloc(seeds.size, realized, durations, cause)
        rc = self._lib.mcdp_ref_run_many(self._h, _pad(seeds), seeds.size, self._p(r, C.c_double),
                                         self._p(d, C.c_double), self._p(c, C.c_int32))
        if rc:
            raise RuntimeError(self._lib.mcdp_ref_last_error().decode())
        return r, d, c


def load_reference_python_module():
    """Import the unmodified reference pybind11 module (``oracle/_ref/_core*.so``) as top-level
    ``_core``.  Its init imports ``mc_dagprop.types`` (reference ``_core.cpp:453``); a stand-in with the
    same NewType names is injected when no ``mc_dagprop`` package is importable."""
    import glob

    paths = glob.glob(os.path.join(_HERE, "_ref", "_core*.so"))
    if not paths:
        raise FileNotFoundError("oracle/_ref/_core*.so missing: run `make -C oracle ref`")
    if "_core" in sys.modules and getattr(sys.modules["_core"], "__file__", "") == paths[0]:
        return sys.modules["_core"]
    try:
        importlib.import_module("mc_dagprop.types")
    except Exception:
        from typing import NewType

        pkg = types.ModuleType("mc_dagprop")
        pkg.__path__ = []
        tm = types.ModuleType("mc_dagprop.types")
        for name, base in [("Second", float), ("ProbabilityMass", float), ("ActivityIndex", int),
                           ("EventIndex", int), ("ActivityType", int), ("EventId", str)]:
            setattr(tm, name, NewType(name, base))
        sys.modules.setdefault("mc_dagprop", pkg)
        sys.modules["mc_dagprop.types"] = tm
    spec = importlib.util.spec_from_file_location("_core", paths[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["_core"] = mod
    return mod
