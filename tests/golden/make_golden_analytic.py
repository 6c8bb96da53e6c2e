"""Generate ``ref_analytic.npz`` from the UNMODIFIED reference analytic engine (pure Python + numpy,
``/root/reference/src/mc_dagprop/analytic``).  Run in the build container:

    python tests/golden/make_golden_analytic.py

The reference package cannot be imported under its own name next to this repository's alias package, and it needs
its compiled ``_core`` plus installed metadata; so the script assembles a private ``mc_dagprop`` namespace from the
reference sources where they lie: ``Event`` / ``EventTimestamp`` from the reference's own pybind11 module built by
``oracle/Makefile`` (oracle/_ref/_core*.so), ``types.py`` and the ``analytic`` sub-package loaded by path.  Outputs
only are stored; ``tests/golden/analytic_cases.py`` regenerates the inputs from the same seeds.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_PKG = "/root/reference/src/mc_dagprop"


def load_reference_analytic():
    sys.path.insert(0, ROOT)
    import oracle  # test infrastructure: the reference builds

    core = oracle.load_reference_python_module()
    sys.path.remove(ROOT)
    pkg = types.ModuleType("mc_dagprop")
    pkg.__path__ = [REF_PKG]
    for name in ("Event", "EventTimestamp", "Activity", "DagContext"):
        setattr(pkg, name, getattr(core, name))
    sys.modules["mc_dagprop"] = pkg

    def load(name, path, is_pkg=False):
        spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)] if is_pkg else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("mc_dagprop.types", os.path.join(REF_PKG, "types.py"))
    analytic = load("mc_dagprop.analytic", os.path.join(REF_PKG, "analytic", "__init__.py"), is_pkg=True)
    ns = types.SimpleNamespace(Event=core.Event, EventTimestamp=core.EventTimestamp, DiscretePMF=analytic.DiscretePMF,
                               AnalyticActivity=analytic.AnalyticActivity, AnalyticContext=analytic.AnalyticContext,
                               UnderflowRule=analytic.UnderflowRule, OverflowRule=analytic.OverflowRule,
                               create_analytic_propagator=analytic.create_analytic_propagator, analytic=analytic)
    return ns


def main():
    ns = load_reference_analytic()
    sys.path.insert(0, HERE)
    import analytic_cases as ac

    out = {}
    for case in ac.CASES:
        name = case[0]
        ctx = ac.build_context(ns, *case)
        try:
            res = ns.create_analytic_propagator(ctx).run()
        except ValueError as exc:
            out[name + "_error"] = np.array([str(exc)])
            continue
        except AssertionError:
            raise SystemExit(f"case {name}: the reference's own mass assertion fired -- not a usable parity case")
        out[name + "_start"] = np.array([r.pmf.values[0] for r in res])
        out[name + "_len"] = np.array([len(r.pmf.values) for r in res])
        out[name + "_probs"] = np.concatenate([r.pmf.probabilities for r in res])
        out[name + "_under"] = np.array([float(r.underflow) for r in res])
        out[name + "_over"] = np.array([float(r.overflow) for r in res])
    for i, (a0, pa, b0, pb) in enumerate(ac.pmf_pairs()):
        a = ns.DiscretePMF(a0 + np.arange(len(pa), dtype=float), pa, step=1)
        b = ns.DiscretePMF(b0 + np.arange(len(pb), dtype=float), pb, step=1)
        c, m = a.convolve(b), a.maximum(b)
        out[f"pair{i}_conv_start"], out[f"pair{i}_conv_probs"] = np.array([c.values[0]]), c.probabilities
        out[f"pair{i}_max_start"], out[f"pair{i}_max_probs"] = np.array([m.values[0]]), m.probabilities
    # the discretised distributions of analytic/distributions.py
    d = ns.analytic
    out["dist_exponential"] = d.exponential_pmf(scale=10.0, step=1, start=0.0, stop=300.0).probabilities
    out["dist_gamma"] = d.gamma_pmf(shape=2.0, scale=2.0, step=1, start=0.0, stop=40.0).probabilities
    out["dist_gamma_half"] = d.gamma_pmf(shape=0.5, scale=3.0, step=2, start=0.0, stop=60.0).probabilities
    out["dist_empirical"] = d.empirical_pmf([0.0, 1.0, 2.0], [1, 1, 2], step=1).probabilities
    path = os.path.join(HERE, "ref_analytic.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", sum(1 for k in out if k.endswith("_error")), "error case(s)")


if __name__ == "__main__":
    main()
