"""Parity at the FULL sizes of the BASELINE.json configurations (``-m gpu``).

For every named configuration the fused kernel runs a batch of seeds on the full-size DAG; its
durations are pushed through the oracle's propagation (reference ``_core.cpp:332-350``) and must
reproduce realized / cause bit-for-bit, and size-independent properties must hold: the
``earliest <= realized <= earliest + max_delay`` envelope, cause consistency, and reduced-mode
statistics equal to the statistics of the full outputs of the same seeds."""
import numpy as np
import pytest

import oracle
from mc_dagprop_b200 import capi, synth

pytestmark = pytest.mark.gpu

CONFIGS = {
    "c1": (synth.c1_toy, 4096),
    "c1_const_exp": (lambda: synth.c1_toy("const_exp"), 4096),
    "c2": (synth.c2_layered, 256),
    "c3": (synth.c3_network, 128),
    "c4": (synth.c4_national, 64),
    "c5": (synth.c5_deep_chain, 128),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_size_config_parity_and_properties(name):
    gen, n = CONFIGS[name]
    dag, dists = gen()
    plan = capi.Plan(dag, dists, device=0)
    osim = oracle.OracleSim(dag, dists)
    seeds = np.arange(1000, 1000 + n, dtype=np.int32)
    r, d, c = plan.run_many_host(seeds)
    assert r.shape == (n, dag.n_events) and d.shape == (n, dag.n_activities)
    # tier 1 at full size: the device's durations through the oracle's propagation
    r_o, c_o = osim.run_injected(d)
    assert np.array_equal(r_o.view(np.uint64), r.view(np.uint64))
    assert np.array_equal(c_o, c)
    # envelope
    assert np.all(r >= dag.earliest[None, :]) and np.all(r <= dag.earliest[None, :] + dag.max_delay)
    # durations never fall below the base duration of a sampled activity (all extras are >= 0 here)
    base = np.zeros(dag.n_activities)
    base[dag.act_idx] = dag.act_base
    assert np.all(d >= base[None, :])
    # cause consistency: an event with a cause is at or after its cause
    rows = np.arange(n)[:, None]
    has = c >= 0
    assert np.all(r[has] >= r[rows.repeat(c.shape[1], 1)[has], c[has]])
    # reduced mode == statistics of the full outputs (integer accumulators exactly)
    delay = r - dag.earliest[None, :]
    th = (60.0, 180.0, 300.0)
    st = plan.run_reduced_host(seeds, thresholds=th, n_bins=64, hist_range=(0.0, dag.max_delay))
    np.testing.assert_allclose(st.sum, delay.sum(0), rtol=1e-12, atol=1e-6)
    np.testing.assert_allclose(st.sumsq, (delay * delay).sum(0), rtol=1e-12, atol=1e-3)
    for i, t in enumerate(th):
        assert np.array_equal(st.late[i], (delay > t).sum(0).astype(np.uint64))
    bins = np.clip(np.floor(delay * (64 / dag.max_delay)).astype(np.int64), 0, 63)
    cnt = np.zeros((dag.n_events, 64), np.int64)
    np.add.at(cnt, (np.arange(dag.n_events)[None, :].repeat(n, 0), bins), 1)
    assert np.array_equal(st.hist.astype(np.int64), cnt)


def test_reduced_mode_many_waves_matches_full_outputs():
    """Several waves of sample groups per launch: every warp stages and folds its own statistics, the global
    accumulators add up to the statistics of the full outputs."""
    dag, dists = synth.random_dag(80, 3, max_delay=60.0), synth.mixed_small_dists()
    plan = capi.Plan(dag, dists, device=0)
    n = 148 * 32 * 64 * 3 + 100  # > 3 batches per warp slot
    seeds = np.arange(n, dtype=np.int32)
    st = plan.run_reduced_host(seeds, thresholds=(1.0, 10.0), n_bins=32, hist_range=(0.0, 60.0))
    acc_sum = np.zeros(plan.E)
    acc_late = np.zeros((2, plan.E), np.uint64)
    acc_hist = np.zeros((plan.E, 32), np.int64)
    for lo in range(0, n, 1 << 17):
        r, _, _ = plan.run_many_host(seeds[lo:lo + (1 << 17)], durations=False, cause=False)
        delay = r - dag.earliest[None, :]
        acc_sum += delay.sum(0)
        for i, t in enumerate((1.0, 10.0)):
            acc_late[i] += (delay > t).sum(0).astype(np.uint64)
        bins = np.clip(np.floor(delay * (32 / 60.0)).astype(np.int64), 0, 31)
        for e in range(plan.E):
            acc_hist[e] += np.bincount(bins[:, e], minlength=32)
    np.testing.assert_allclose(st.sum, acc_sum, rtol=1e-11)
    assert np.array_equal(st.late, acc_late) and np.array_equal(st.hist.astype(np.int64), acc_hist)
    assert int(st.hist.sum()) == n * plan.E
