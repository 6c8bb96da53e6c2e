"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container (needs oracle/_ref, i.e. /root/reference):

    python tests/golden/make_golden.py

* ``ref_random_dags.npz``  -- outputs of the reference C++ engine (oracle/_ref/libmcdp_ref.so) for
  seeded random DAGs (``mc_dagprop_b200.synth.random_dag``) with every distribution kind.
* ``ref_python_api.npz``   -- outputs of the reference's own pybind11 module (oracle/_ref/_core*.so)
  driven through its Python API on the fixtures of reference test/test_simulator.py and
  test/test_monte_carlo_extra.py.
The fixtures hold outputs only; inputs are regenerated from the same seeds by the tests.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from mc_dagprop_b200 import synth  # noqa: E402

CASES = [(5, 100), (40, 101), (150, 102), (400, 103)]  # (n_events, dag seed)
SEEDS = np.arange(-3, 21, dtype=np.int32)


def random_dag_cases():
    out = {}
    for n, s in CASES:
        for md in (50.0, 4.0):
            dag = synth.random_dag(n, s, max_delay=md)
            ref = oracle.RefSim(dag, synth.mixed_small_dists())
            r, d, c = ref.run_many(SEEDS)
            r2, d2, c2 = ref.run_many(SEEDS)  # second pass: gamma's cached normal leaks across runs
            key = f"n{n}_s{s}_md{int(md)}"
            out[key + "_realized"], out[key + "_durations"], out[key + "_cause"] = r, d, c
            out[key + "_realized2"], out[key + "_durations2"], out[key + "_cause2"] = r2, d2, c2
    return out


def python_api_cases():
    m = oracle.load_reference_python_module()
    EventTimestamp, Event, Activity, DagContext = m.EventTimestamp, m.Event, m.Activity, m.DagContext
    events = [Event(str(i), EventTimestamp(e, 100.0, 0.0)) for i, e in enumerate([0.0, 5.0, 10.0, 22.0, 20.0, 100.0])]
    link_map = {(0, 1): Activity(0, 3.0, 1), (1, 2): Activity(1, 5.0, 1), (1, 3): Activity(2, 5.0, 1),
                (2, 4): Activity(3, 15.0, 2), (3, 4): Activity(4, 10.0, 3)}
    prec = [(1, [(0, 0)]), (2, [(1, 1)]), (3, [(1, 2)]), (4, [(2, 3), (3, 4)])]
    ctx = DagContext(events=events, activities=link_map, precedence_list=prec, max_delay=1e6)
    out = {}

    def record(name, gen, seeds):
        sim = m.MonteCarloPropagator(ctx, gen)
        res = sim.run_many(list(seeds))
        out[name + "_realized"] = np.array([r.realized for r in res])
        out[name + "_durations"] = np.array([r.durations for r in res])
        out[name + "_cause"] = np.array([r.cause_event for r in res])

    g = m.GenericDelayGenerator()
    g.add_constant(activity_type=1, factor=1.0)
    g.add_constant(activity_type=3, factor=3.0)
    record("constant", g, range(5))
    g = m.GenericDelayGenerator()
    g.add_empirical_absolute(activity_type=1, values=[10, 20, 40, 50], weights=[0.1, 0.2, 0.3, 0.4])
    record("emp_abs", g, range(16))
    g = m.GenericDelayGenerator()
    g.add_empirical_relative(activity_type=1, factors=[1.2, 1.3, 1.35, 4.5], weights=[0.1, 0.2, 0.3, 0.4])
    record("emp_rel", g, range(16))
    g = m.GenericDelayGenerator()
    g.add_exponential(1, 1000.0, max_scale=1.0)
    record("exp_rejection_heavy", g, range(3))
    g = m.GenericDelayGenerator()
    g.add_gamma(activity_type=1, shape=2.0, scale=1.0, max_scale=0.5)
    record("gamma_truncated", g, range(8))
    return out


if __name__ == "__main__":
    if not oracle.have_ref():
        raise SystemExit("oracle/_ref missing: run `make -C oracle ref` first (needs /root/reference)")
    np.savez_compressed(os.path.join(HERE, "ref_random_dags.npz"), **random_dag_cases())
    np.savez_compressed(os.path.join(HERE, "ref_python_api.npz"), **python_api_cases())
    for f in ("ref_random_dags.npz", "ref_python_api.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
