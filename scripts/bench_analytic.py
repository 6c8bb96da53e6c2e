"""Timing of the analytic PMF propagator on one layered random DAG.

    python scripts/bench_analytic.py [--events 2000] [--reps 5]            # drop-in package, GPU engine (B200 box)
    python scripts/bench_analytic.py --impl reference [--events 2000]      # the reference's numpy engine (build container only:
                                                                           # it is imported from /root/reference, which does not travel)

Both arms build the same AnalyticContext from tests/golden/analytic_cases.describe (wide windows, supports of up to 300
bins per edge, TRUNCATE rules) and time ``create_analytic_propagator(ctx)`` and ``propagator.run()`` separately.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--seed", type=int, default=21)
    args = ap.parse_args()
    if args.impl == "reference":
        import make_golden_analytic as mg

        ns = mg.load_reference_analytic()
        create = ns.create_analytic_propagator
    else:
        sys.path.insert(0, ROOT)
        if os.environ.get("MCDP_LIB"):  # scratch A/B builds (scripts/ab_variants.sh)
            from mc_dagprop_b200 import capi

            capi.LIB_PATH = os.path.abspath(os.environ["MCDP_LIB"])
        import mc_dagprop
        import mc_dagprop.analytic as an

        ns = argparse.Namespace(Event=mc_dagprop.Event, EventTimestamp=mc_dagprop.EventTimestamp, DiscretePMF=an.DiscretePMF,
                                AnalyticActivity=an.AnalyticActivity, AnalyticContext=an.AnalyticContext,
                                UnderflowRule=an.UnderflowRule, OverflowRule=an.OverflowRule)
        create = an.create_analytic_propagator
    import analytic_cases as ac

    t0 = time.perf_counter()
    ctx = ac.build_context(ns, "wide_bench", args.events, args.seed, 1, None, 1, 1)
    t_ctx = time.perf_counter() - t0
    t0 = time.perf_counter()
    prop = create(ctx)
    t_create = time.perf_counter() - t0
    out = prop.run()  # warm-up (device buffers, module load)
    times = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        out = prop.run()
        times.append(time.perf_counter() - t0)
    profile = None
    if args.impl == "b200":
        from mc_dagprop_b200.analytic import _device

        profile = _device.last_profile()  # of the last repetition
    bins_in = sum(len(a.pmf.probabilities) for _, a in ctx.activities.values())
    bins_out = sum(len(e.pmf.probabilities) for e in out)
    print(json.dumps({"impl": args.impl, "workload": "analytic: layered random DAG, wide windows, TRUNCATE rules",
                      "events": len(ctx.events), "activities": len(ctx.activities), "input_pmf_bins": bins_in,
                      "output_pmf_bins": bins_out, "context_build_s": t_ctx, "create_propagator_s": t_create,
                      "run_s_best": min(times), "run_s_median": sorted(times)[len(times) // 2], "reps": args.reps,
                      "events_per_s": len(ctx.events) / min(times), "library_call_phases": profile}))


if __name__ == "__main__":
    main()
