"""Timings of the rows SURVEY.md section 8(f) lists either side of the hot path, on configuration 3 (100k events /
399k activities), each beside the reference where the reference has the step:

  f-1  ingest: list[Event] + dict[(src, dst) -> Activity] + precedence_list -> DagContext -> propagator (the reference's
       only way in, _core.cpp:425-428, 193-305) against MonteCarloPropagator.from_arrays;
  f-2  the reference-compatible generator stream (MCDP_OPT_RNG_STREAM = 1) against the default Philox stream and
       against the reference's C++ engine on one host thread;
  f-3  device-side consumers: reduced statistics and delay-cause attribution against full outputs + numpy.

    python scripts/bench_next_rows.py            # on a B200 box (the reference module travels in oracle/_ref)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fill_generator(gen, dists):
    for t, (kind, a, b, c, v, w) in dists._entries.items():
        if kind == 0:
            gen.add_constant(t, a)
        elif kind == 1:
            gen.add_exponential(t, a, b)
        elif kind == 2:
            gen.add_gamma(t, a, b, c)
        elif kind == 3:
            gen.add_empirical_absolute(t, list(v), list(w))
        else:
            gen.add_empirical_relative(t, list(v), list(w))
    return gen


def as_objects(mod, dag):
    """The reference's input format from the flat arrays: Python objects, one per event and per activity."""
    events = [mod.Event(str(i), mod.EventTimestamp(float(e), float(e) + 1e6, float(e))) for i, e in enumerate(dag.earliest)]
    pos = {int(a): k for k, a in enumerate(dag.act_idx)}
    activities, precedence = {}, []
    for i, t in enumerate(dag.prec_target):
        preds = []
        for k in range(int(dag.prec_off[i]), int(dag.prec_off[i + 1])):
            s, a = int(dag.pred_src[k]), int(dag.pred_act[k])
            j = pos[a]
            activities[(s, int(t))] = mod.Activity(a, float(dag.act_base[j]), int(dag.act_type[j]))
            preds.append((s, a))
        precedence.append((int(t), preds))
    return events, activities, precedence


def timed(fn, reps=1):
    best = float("inf")
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    import mc_dagprop
    import oracle
    from bench import THRESHOLDS, N_BINS, build_workload

    dag, dists = build_workload("c3")
    E, A = dag.n_events, dag.n_activities
    res = {"workload": "c3", "events": E, "activities": A}

    # ---- f-1 ingest ------------------------------------------------------------------------------------------
    ingest = {}
    gen = fill_generator(mc_dagprop.GenericDelayGenerator(), dists)
    from_arrays = lambda d=dag: mc_dagprop.MonteCarloPropagator.from_arrays(  # noqa: E731
        d.earliest, d.act_idx, d.act_base, d.act_type, d.prec_target, d.prec_off, d.pred_src, d.pred_act, d.max_delay, gen)
    t_first, sim = timed(from_arrays)  # first plan of the process: CUDA context, module load
    ingest["from_arrays_first_call_s"] = t_first
    t, sim = timed(from_arrays, reps=3)
    ingest["from_arrays_s"] = t
    t_obj, objs = timed(lambda: as_objects(mc_dagprop, dag))
    ingest["python_objects_s"] = t_obj
    t_ctx, ctx = timed(lambda: mc_dagprop.DagContext(objs[0], objs[1], objs[2], dag.max_delay))
    t_sim, sim_o = timed(lambda: mc_dagprop.MonteCarloPropagator(ctx, gen), reps=3)
    ingest["dagcontext_s"], ingest["propagator_from_context_s"] = t_ctx, t_sim
    # the dict keeps one activity per (src, dst) pair (627 parallel edges of the synthetic DAG collapse, their indices
    # become zero-duration links, _core.cpp:213-231): compare with the array path on exactly that activity set
    from mc_dagprop_b200.flat import FlatDag

    kept = list(objs[1].values())
    same = FlatDag(dag.earliest, [a.idx for a in kept], [a.minimal_duration for a in kept], [a.activity_type for a in kept],
                   dag.prec_target, dag.prec_off, dag.pred_src, dag.pred_act, dag.max_delay)
    a = from_arrays(same).run_many_arrays(np.arange(4, dtype=np.int32))
    b = sim_o.run_many_arrays(np.arange(4, dtype=np.int32))
    ingest["same_results_both_ways"] = bool(all(np.array_equal(x, y) for x, y in zip(a, b)))
    del sim_o, ctx, objs
    if oracle.have_ref():
        ref = oracle.load_reference_python_module()
        rgen = fill_generator(ref.GenericDelayGenerator(), dists)
        t_obj, robjs = timed(lambda: as_objects(ref, dag))
        t_ctx, rctx = timed(lambda: ref.DagContext(robjs[0], robjs[1], robjs[2], dag.max_delay))
        t_sim, rsim = timed(lambda: ref.MonteCarloPropagator(rctx, rgen), reps=3)
        ingest["reference"] = {"python_objects_s": t_obj, "dagcontext_s": t_ctx, "simulator_s": t_sim}
        del rsim, rctx, robjs
    res["f1_ingest"] = ingest

    # ---- f-2 reference-compatible stream ------------------------------------------------------------------------
    n = 2048
    seeds = np.arange(n, dtype=np.int32)
    compat = {"samples": n}
    sim.run_many_arrays(seeds)
    t, _ = timed(lambda: sim.run_many_arrays(seeds), reps=3)
    compat["philox_stream_s"] = t
    sim.set_option(4, 1)
    sim.run_many_arrays(seeds)
    t, _ = timed(lambda: sim.run_many_arrays(seeds), reps=3)
    compat["reference_stream_s"] = t
    sim.set_option(4, 0)
    if oracle.have_ref():
        rs = oracle.RefSim(dag, dists)
        k = 64
        t, _ = timed(lambda: rs.run_many(seeds[:k]))
        compat["reference_cpp_one_thread_s_per_sample"] = t / k
        compat["reference_cpp_one_thread_s_for_all"] = t / k * n
    compat["edge_samples_per_s_reference_stream"] = n * A / compat["reference_stream_s"]
    res["f2_reference_stream"] = compat

    # ---- f-3 device-side consumers ------------------------------------------------------------------------------
    n = 18944
    seeds = np.arange(n, dtype=np.int32)
    cons = {"samples": n}
    kw = dict(thresholds=list(THRESHOLDS), n_bins=N_BINS, hist_lo=0.0, hist_hi=float(dag.max_delay))
    sim.run_many_reduced(seeds, **kw)
    t, st = timed(lambda: sim.run_many_reduced(seeds, **kw), reps=3)
    cons["reduced_statistics_s"] = t
    sim.run_many_reduced(seeds, cause_counts=True, **kw)
    t, st2 = timed(lambda: sim.run_many_reduced(seeds, cause_counts=True, **kw), reps=3)
    cons["reduced_with_cause_attribution_s"] = t
    m = 1024  # the same statistics the way a user of the reference gets them: full outputs to the host, numpy
    sim.run_many_arrays(seeds[:m])

    def numpy_way():
        r, d, c = sim.run_many_arrays(seeds[:m])
        delay = r - np.asarray(dag.earliest)[None, :]
        s1, s2 = delay.sum(axis=0), (delay * delay).sum(axis=0)
        late = [(delay > th).sum(axis=0) for th in THRESHOLDS]
        return s1, s2, late

    t, _ = timed(numpy_way)
    cons["full_outputs_plus_numpy_s_per_sample"] = t / m
    cons["full_outputs_plus_numpy_s_for_all"] = t / m * n
    cons["edge_samples_per_s_reduced"] = n * A / cons["reduced_statistics_s"]
    res["f3_consumers"] = cons
    print(json.dumps(res))


if __name__ == "__main__":
    main()
