"""Several devices behind one call (``mcdp_planset_*`` / ``MonteCarloPropagator(..., devices=[...])``): a call
over a plan set must return exactly what a single device returns -- identical bits for full outputs, duration
injection and the integer statistics, f64 sums to summation order.  On a single-GPU box the set lists device 0
more than once (two plans, two streams, the same sharding and the same peer reduction); with two or more GPUs the
same checks run across devices (seed sharding equivalence, SURVEY.md section 4 implication 2)."""
import numpy as np
import pytest

from mc_dagprop_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _device_lists():
    n = capi.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists += [[0, 1], list(range(min(n, 8)))]
    return lists


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.mark.parametrize("n", [0, 1, 127, 128, 129, 700, 5000])
def test_full_outputs_identical_to_one_device(n):
    dag, dists = synth.random_dag(400, 21), synth.mixed_small_dists()
    one = capi.Plan(dag, dists, device=0)
    seeds = np.arange(-3, n - 3, dtype=np.int32)
    ref = one.run_many_host(seeds)
    for devs in _device_lists():
        ps = capi.PlanSet(dag, dists, devs)
        assert len(ps) == len(devs)
        got = ps.run_many_host(seeds)
        for a, b in zip(ref, got):
            assert a.shape == b.shape and np.array_equal(_bits(a), _bits(b)), devs
        if n:
            r_i, c_i = ps.run_injected_host(ref[1])
            assert np.array_equal(_bits(r_i), _bits(ref[0])) and np.array_equal(c_i, ref[2])


@pytest.mark.parametrize("n", [0, 5, 1000, 20000])
def test_reduced_statistics_equal_one_device(n):
    dag, dists = synth.random_dag(600, 22, max_delay=60.0), synth.mixed_small_dists()
    one = capi.Plan(dag, dists, device=0)
    seeds = np.arange(100, 100 + n, dtype=np.int32)
    th = (1.0, 10.0, 30.0)
    ref, ref_act, ref_none = one.run_attribution_host(seeds, thresholds=th, n_bins=24, hist_range=(0.0, 60.0))
    for devs in _device_lists():
        ps = capi.PlanSet(dag, dists, devs)
        st, act, none = ps.run_attribution_host(seeds, thresholds=th, n_bins=24, hist_range=(0.0, 60.0))
        np.testing.assert_allclose(st.sum, ref.sum, rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(st.sumsq, ref.sumsq, rtol=1e-12, atol=1e-9)
        assert np.array_equal(st.late, ref.late) and np.array_equal(st.hist, ref.hist), devs
        assert np.array_equal(act, ref_act) and np.array_equal(none, ref_none), devs
        st2 = ps.run_reduced_host(seeds, thresholds=th, n_bins=24, hist_range=(0.0, 60.0))  # repeatable, no stale state
        assert np.array_equal(st2.hist, ref.hist) and np.array_equal(st2.late, ref.late)
        np.testing.assert_allclose(st2.sum, st.sum, rtol=1e-12, atol=1e-9)  # (f64 atomics: order varies run to run)
        only_hist = ps.run_reduced_host(seeds, n_bins=24, hist_range=(0.0, 60.0))
        assert only_hist.late.shape == (0, ps.E) and np.array_equal(only_hist.hist, ref.hist)


def test_options_reach_every_plan_and_errors_name_the_device():
    dag, dists = synth.random_dag(200, 23), synth.mixed_small_dists()
    ps = capi.PlanSet(dag, dists, [0, 0])
    ps.set_option(capi.OPT_STREAM_KEY, 77)
    one = capi.Plan(dag, dists, device=0)
    one.set_option(capi.OPT_STREAM_KEY, 77)
    seeds = np.arange(300, dtype=np.int32)
    assert np.array_equal(_bits(ps.run_many_host(seeds)[1]), _bits(one.run_many_host(seeds)[1]))
    with pytest.raises(RuntimeError, match="samples per lane"):
        ps.set_option(capi.OPT_SAMPLES_PER_LANE, 3)
    with pytest.raises(RuntimeError, match="1..16 devices"):
        capi.PlanSet(dag, dists, [])
    with pytest.raises(RuntimeError, match="out of range"):
        capi.PlanSet(dag, dists, [0, 99])


def test_drop_in_propagator_over_several_devices():
    """The reference-facing class: ``devices=[...]`` changes nothing but where the samples are computed."""
    from mc_dagprop import Activity, DagContext, Event, EventTimestamp, GenericDelayGenerator, MonteCarloPropagator

    rng = np.random.default_rng(5)
    n_ev = 60
    events = [Event(str(i), EventTimestamp(float(3 * i), 1e9, 0.0)) for i in range(n_ev)]
    acts, prec, idx = {}, [], 0
    for tgt in range(1, n_ev):
        preds = []
        for src in sorted(set(rng.integers(0, tgt, size=2).tolist())):
            acts[(src, tgt)] = Activity(idx=idx, minimal_duration=float(rng.integers(1, 6)), activity_type=1 + idx % 3)
            preds.append((src, idx))
            idx += 1
        prec.append((tgt, preds))
    ctx = DagContext(events=events, activities=acts, precedence_list=prec, max_delay=50.0)
    gen = GenericDelayGenerator()
    gen.add_exponential(1, 0.5, 3.0)
    gen.add_gamma(2, 2.0, 0.3, 4.0)
    gen.add_empirical_relative(3, [0.0, 0.5, 1.0, 2.0], [0.4, 0.3, 0.2, 0.1])
    one = MonteCarloPropagator(ctx, gen)
    many = MonteCarloPropagator(ctx, gen, devices=[0, 0])
    assert many.devices() == [0, 0] and one.devices() == [0] and many.device() == 0
    seeds = list(range(-10, 400))
    a, b = one.run_many(seeds), many.run_many(seeds)
    assert len(a) == len(b) == len(seeds)
    for x, y in zip(a, b):
        assert np.array_equal(x.realized, y.realized) and np.array_equal(x.durations, y.durations)
        assert np.array_equal(x.cause_event, y.cause_event)
    ra = one.run_many_reduced(np.asarray(seeds, np.int32), [1.0], 8, 0.0, 50.0, True)
    rb = many.run_many_reduced(np.asarray(seeds, np.int32), [1.0], 8, 0.0, 50.0, True)
    for k in ("late", "hist", "cause_activity", "cause_none"):
        assert np.array_equal(ra[k], rb[k]), k
    np.testing.assert_allclose(ra["sum"], rb["sum"], rtol=1e-12, atol=1e-9)
    r1 = many.run(seed=7)
    assert np.array_equal(r1.realized, one.run(seed=7).realized) and r1.realized.base is r1
