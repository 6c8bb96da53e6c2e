"""Calls of a handful of samples (``run(seed)``, reference ``_core.cpp:312-353``) take a path of their own
(``csrc/mcdp_small_sweep.cuh``: one thread per activity draws, one thread per event propagates).  It has to return
the bits of the sweep kernels: same generator contract, same order of the recurrence."""
import numpy as np
import pytest

from mc_dagprop_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _both(plan, call):
    plan.set_option(capi.OPT_SMALL_CALL_MAX, 0)
    sweep = call()
    plan.set_option(capi.OPT_SMALL_CALL_MAX, 4096)
    small = call()
    plan.set_option(capi.OPT_SMALL_CALL_MAX, -1)
    return sweep, small


def _same(a, b):
    return all(np.array_equal(x.view(np.uint64) if x.dtype == np.float64 else x, y.view(np.uint64) if y.dtype == np.float64 else y)
               for x, y in zip(a, b))


@pytest.mark.parametrize("seed", [3, 17, 29])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 33, 64, 130])
def test_fused_sampling_matches_the_sweep_kernels(seed, n):
    """Every distribution kind (both gamma samplers, exponentials incl. the series branch, narrow and wide tables,
    constants, untyped activities), index gaps, orphan activities, arbitrary seeds."""
    dag = synth.random_dag(400, seed=seed, idx_gaps=True)
    plan = capi.Plan(dag, synth.mixed_small_dists(), device=0)
    rng = np.random.default_rng(seed)
    seeds = rng.integers(-2**31, 2**31 - 1, size=n).astype(np.int32)
    sweep, small = _both(plan, lambda: plan.run_many_host(seeds))
    assert plan.launch_shape(1)["samples_per_lane"] == 1
    assert _same(sweep, small)
    seq = np.arange(7, 7 + n, dtype=np.int32)  # consecutive seeds: the sweep kernels share blocks between neighbours
    sweep, small = _both(plan, lambda: plan.run_many_host(seq))
    assert _same(sweep, small)


def test_generic_gamma_and_large_tables():
    from mc_dagprop_b200.flat import FlatDists

    dag = synth.random_dag(300, seed=5, n_types=5)
    d = FlatDists()
    d.add_gamma(0, 2.3, 0.1, 5.0)      # Marsaglia-Tsang
    d.add_gamma(1, 0.6, 0.3, 4.0)      # shape < 1: boost
    d.add_gamma(2, 2.0, 0.2, 0.05)     # tight bound: retries up to the cap
    x = np.linspace(0.0, 3.0, 6000)    # more than 4096 entries: 64-bit draws
    d.add_empirical_relative(3, x, np.exp(-x))
    d.add_exponential(4, 0.4, 1e-4)    # tiny F: series branch
    plan = capi.Plan(dag, d, device=0)
    for n in (1, 4, 31):
        seeds = np.arange(-5, -5 + n, dtype=np.int32)
        sweep, small = _both(plan, lambda: plan.run_many_host(seeds))
        assert _same(sweep, small)


def test_injected_durations_and_reference_stream():
    import oracle

    dag = synth.random_dag(500, seed=11, idx_gaps=True)
    dists = synth.mixed_small_dists()
    plan = capi.Plan(dag, dists, device=0)
    seeds = np.arange(6, dtype=np.int32)
    _, dur, _ = oracle.OracleSim(dag, dists).run_many(seeds)
    dur[1, ::7] = np.inf
    dur[2, ::5] = -3.5
    dur[3, 3] = np.nan
    sweep, small = _both(plan, lambda: plan.run_injected_host(dur))
    assert _same(sweep, small)
    r_o, c_o = oracle.OracleSim(dag, dists).run_injected(dur)  # and both are the reference's propagation
    assert np.array_equal(r_o.view(np.uint64), small[0].view(np.uint64)) and np.array_equal(c_o, small[1])
    plan.set_option(capi.OPT_RNG_STREAM, 1)
    sweep, small = _both(plan, lambda: plan.run_many_host(seeds))
    assert _same(sweep, small)


def test_full_size_network_single_seed():
    dag, dists = synth.c3_network()
    plan = capi.Plan(dag, dists, device=0)
    seeds = np.array([20261003], dtype=np.int32)
    sweep, small = _both(plan, lambda: plan.run_many_host(seeds))
    assert _same(sweep, small)
    assert plan.launch_shape(1)["samples_per_lane"] == 1 and plan.launch_shape(18944)["samples_per_lane"] == 4


def test_drop_in_run_is_a_row_of_run_many():
    """reference test_monte_carlo_extra.py:101-109 through the drop-in: run(seed) (small-call path) against the same
    seed inside a large batch (sweep kernels)."""
    from mc_dagprop import GenericDelayGenerator, MonteCarloPropagator

    dag = synth.random_dag(300, seed=2)
    gen = GenericDelayGenerator()
    gen.add_gamma(0, 2.0, 0.1, 5.0)
    gen.add_gamma(1, 2.3, 0.2, 5.0)
    gen.add_exponential(2, 0.3, 4.0)
    gen.add_empirical_relative(3, [0.0, 0.5, 1.0], [2.0, 1.0, 1.0])
    gen.add_empirical_absolute(4, [0.0, 10.0], [3.0, 1.0])
    gen.add_constant(5, 0.25)
    prop = MonteCarloPropagator.from_arrays(dag.earliest, dag.act_idx, dag.act_base, dag.act_type, dag.prec_target, dag.prec_off,
                                            dag.pred_src, dag.pred_act, dag.max_delay, gen)
    r, d, c = prop.run_many_arrays(np.arange(1000, dtype=np.int32))
    for seed in (0, 1, 499, 999):
        one = prop.run(seed=seed)
        assert np.array_equal(one.realized, r[seed]) and np.array_equal(one.durations, d[seed]) and np.array_equal(one.cause_event, c[seed])
