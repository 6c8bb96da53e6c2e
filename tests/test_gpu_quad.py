"""GPU parity tests of the quad sweep kernel (four samples per lane, ``MCDP_OPT_SAMPLES_PER_LANE = 4``,
mc_dagprop_b200/csrc/mcdp_quad_sweep.cuh), called through the C ABI.

The quad kernel must return exactly what the pair kernel returns (same generator contract, same recurrence):
  * duration injection: realized bit-exact in fp64, cause_event exact against the CPU oracle;
  * fused sampling: durations / realized / cause bit-identical to the pair kernel for paired, unpaired,
    arbitrary and extreme seeds, for every distribution kind (Erlang-shape and Marsaglia-Tsang gamma included);
  * reduced statistics and delay-cause attribution: integer accumulators exact, fp64 sums to summation order;
  * ragged sample counts, row lengths that are a multiple of 64 but not of 128 (shadow lanes), buffers that
    are only 16-byte aligned (falls back to the pair kernel).
"""
import numpy as np
import pytest

import oracle
from mc_dagprop_b200 import capi, synth
from mc_dagprop_b200.flat import FlatDists

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _plan(dag, dists, spl, wpg=0, gpc=0, chunk=0):
    plan = capi.Plan(dag, dists, device=0)
    plan.set_option(capi.OPT_SAMPLES_PER_LANE, spl)
    if wpg:
        plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
    if gpc:
        plan.set_option(capi.OPT_GROUPS_PER_CTA, gpc)
    if chunk:
        plan.set_option(capi.OPT_HOST_CHUNK, chunk)
    return plan


def _all_kinds_dists() -> FlatDists:
    """mixed_small_dists plus every Erlang-shape gamma variant and a truncating max_scale."""
    d = synth.mixed_small_dists()
    for t, shape in enumerate((0.5, 1.0, 1.5, 2.0, 3.0, 3.5, 4.0), start=7):
        d.add_gamma(t, shape, 0.3, 0.8 if t % 2 else 5.0)
    return d


@pytest.mark.parametrize("n_events,seed", [(2, 0), (17, 1), (200, 2), (1500, 3)])
@pytest.mark.parametrize("wpg", [1, 3, 8, 16])
def test_quad_injected_bit_exact_random_dags(n_events, seed, wpg):
    dag = synth.random_dag(n_events, seed)
    dists = synth.mixed_small_dists()
    osim = oracle.OracleSim(dag, dists)
    _, dur, _ = osim.run_many(np.arange(-7, 150, dtype=np.int32))  # 157 samples: ld = 192, shadow lanes in group 1
    r_o, c_o = osim.run_injected(dur)
    r_d, c_d = _plan(dag, dists, 4, wpg=wpg).run_injected_host(dur)
    assert np.array_equal(_bits(r_o), _bits(r_d))
    assert np.array_equal(c_o, c_d)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 127, 128, 129, 200, 999])
def test_quad_injected_ragged_counts(n):
    dag = synth.random_dag(150, 9, max_delay=25.0)
    dists = synth.mixed_small_dists()
    osim = oracle.OracleSim(dag, dists)
    _, dur, _ = osim.run_many(np.arange(n, dtype=np.int32))
    r_o, c_o = osim.run_injected(dur)
    for chunk in (0, 192):
        r_d, c_d = _plan(dag, dists, 4, chunk=chunk).run_injected_host(dur)
        assert np.array_equal(_bits(r_o), _bits(r_d))
        assert np.array_equal(c_o, c_d)


@pytest.mark.parametrize("wpg", [0, 1, 4, 16])
def test_quad_matches_pair(wpg):
    """Fused sampling + sweep: bit-identical outputs from both kernels, self-consistent through the oracle."""
    dag = synth.random_dag(400, 20, n_types=14)
    dists = _all_kinds_dists()
    rng = np.random.default_rng(wpg)
    # aligned quads, odd start (no pairs), pairs that are not quads, arbitrary, extremes
    seeds = np.concatenate([np.arange(0, 260), np.arange(1001, 1100), np.arange(2, 80), rng.integers(-2**31, 2**31 - 1, size=100),
                            [-1, 2**31 - 1, -2**31]]).astype(np.int32)
    r2, d2, c2 = _plan(dag, dists, 2, wpg=wpg).run_many_host(seeds)
    r4, d4, c4 = _plan(dag, dists, 4, wpg=wpg).run_many_host(seeds)
    assert np.array_equal(_bits(d2), _bits(d4))
    assert np.array_equal(_bits(r2), _bits(r4))
    assert np.array_equal(c2, c4)
    osim = oracle.OracleSim(dag, dists)
    ro, co = osim.run_injected(d4)
    assert np.array_equal(_bits(ro), _bits(r4)) and np.array_equal(co, c4)
    # a sample is a pure function of its seed: permuting the seeds changes nothing
    perm = rng.permutation(seeds.size)
    r5, d5, c5 = _plan(dag, dists, 4, wpg=wpg, chunk=256).run_many_host(seeds[perm])
    assert np.array_equal(_bits(d5), _bits(d4[perm])) and np.array_equal(_bits(r5), _bits(r4[perm]))
    assert np.array_equal(c5, c4[perm])


def test_quad_matches_generator_contract():
    dag = synth.random_dag(300, 11)
    dists = synth.mixed_small_dists()
    osim = oracle.OracleSim(dag, dists)
    seeds = np.arange(-3, 253, dtype=np.int32)
    _, d, _ = _plan(dag, dists, 4).run_many_host(seeds)
    _, d_spec, _ = osim.run_many_spec(seeds)
    kinds = {t: k for t, k in zip(dists.dist_type.tolist(), dists.kind.tolist())}
    act_kind = np.full(osim.A, -1)
    for i, t in zip(dag.act_idx.tolist(), dag.act_type.tolist()):
        act_kind[i] = kinds.get(t, -1)
    exact = np.isin(act_kind, [-1, 0, 3, 4])
    assert np.array_equal(_bits(d[:, exact]), _bits(d_spec[:, exact]))
    expo = act_kind == 1
    np.testing.assert_allclose(d[:, expo], d_spec[:, expo], rtol=1e-9, atol=1e-12)
    gam = act_kind == 2
    assert np.isclose(d[:, gam], d_spec[:, gam], rtol=5e-5, atol=1e-9).mean() > 0.9995


@pytest.mark.parametrize("wpg", [1, 4])
def test_quad_reduced_and_attribution_match_pair(wpg):
    dag = synth.random_dag(500, 41, max_delay=40.0)
    dists = synth.mixed_small_dists()
    seeds = np.arange(3, 3 + 2001, dtype=np.int32)
    th = (1.0, 5.0, 20.0)
    kw = dict(thresholds=th, n_bins=16, hist_range=(0.0, 40.0))
    p2, p4 = _plan(dag, dists, 2, wpg=wpg), _plan(dag, dists, 4, wpg=wpg)
    s2, s4 = p2.run_reduced_host(seeds, **kw), p4.run_reduced_host(seeds, **kw)
    assert np.array_equal(s2.hist, s4.hist) and np.array_equal(s2.late, s4.late)
    assert int(s4.hist.sum()) == seeds.size * p4.E
    np.testing.assert_allclose(s4.sum, s2.sum, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(s4.sumsq, s2.sumsq, rtol=1e-12, atol=1e-9)
    # against the full outputs
    r, _, c = p4.run_many_host(seeds)
    delay = r - dag.earliest[None, :]
    for i, t in enumerate(th):
        assert np.array_equal(s4.late[i], (delay > t).sum(0).astype(np.uint64))
    (a2, act2, none2), (a4, act4, none4) = p2.run_attribution_host(seeds, **kw), p4.run_attribution_host(seeds, **kw)
    assert np.array_equal(act2, act4) and np.array_equal(none2, none4) and np.array_equal(a2.hist, a4.hist)
    assert np.array_equal(none4, (c == -1).sum(0).astype(np.uint64))


@pytest.mark.parametrize("seed0", [0, 4096, 7, -6, 2])
@pytest.mark.parametrize("ld,offset", [(320, 0), (384, 0), (320, 2), (384, 2)])
def test_quad_device_call_layouts(seed0, ld, offset):
    """Device-buffer calls: row lengths 64 * odd (shadow lanes) and 128-multiples; buffers offset by 16 bytes are not
    32-byte aligned, so the call takes the pair kernel -- the result must not depend on any of it."""
    import torch

    dag, dists = synth.random_dag(400, 21), synth.mixed_small_dists()
    plan = _plan(dag, dists, 4)
    n = 300
    E, A = plan.E, plan.A
    dev = torch.device("cuda:0")
    rb = torch.zeros(E * ld + 4, dtype=torch.float64, device=dev)
    db = torch.zeros(A * ld + 4, dtype=torch.float64, device=dev)
    cb = torch.zeros(E * ld + 8, dtype=torch.int32, device=dev)
    r = rb[offset:offset + E * ld].view(E, ld)
    d = db[offset:offset + A * ld].view(A, ld)
    c = cb[2 * offset:2 * offset + E * ld].view(E, ld)
    plan.run_full_device(n, r, d, c, ld, seed0=seed0)
    torch.cuda.synchronize()
    seeds = np.arange(seed0, seed0 + n, dtype=np.int32)
    r_h, d_h, c_h = _plan(dag, dists, 2).run_many_host(seeds)
    assert np.array_equal(_bits(r[:, :n].T.cpu().numpy()), _bits(r_h))
    assert np.array_equal(_bits(d[:, :n].T.cpu().numpy()), _bits(d_h))
    assert np.array_equal(c[:, :n].T.cpu().numpy(), c_h)
    # nothing outside the row range was touched
    assert float(rb[:offset].abs().sum()) == 0.0 and float(rb[offset + E * ld:].abs().sum()) == 0.0
    assert float(db[:offset].abs().sum()) == 0.0 and float(db[offset + A * ld:].abs().sum()) == 0.0
    # injected through device buffers
    r2 = torch.zeros_like(r)
    c2 = torch.zeros_like(c)
    plan.run_injected_device(n, d, r2, c2, ld)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(r2[:, :n].T.cpu().numpy()), _bits(r_h))
    assert np.array_equal(c2[:, :n].T.cpu().numpy(), c_h)


def test_quad_full_size_c3_slice_matches_pair():
    """The bench workload's DAG (100k events / 399k activities) on one 128-sample group per kernel."""
    dag, dists = synth.c3_network()
    seeds = np.arange(1000, 1000 + 192, dtype=np.int32)
    r2, d2, c2 = _plan(dag, dists, 2).run_many_host(seeds)
    r4, d4, c4 = _plan(dag, dists, 4).run_many_host(seeds)
    assert np.array_equal(_bits(d2), _bits(d4)) and np.array_equal(_bits(r2), _bits(r4)) and np.array_equal(c2, c4)


@pytest.mark.parametrize("name,n", [("c2", 32768), ("c3", 18944)])
def test_quad_machine_filling_launch_matches_oracle_and_pair(name, n):
    """The launch shapes the benchmark times (auto choice = quad kernel, every SM busy: C2 in 256 groups of 10 warps, two
    CTAs per SM; C3 at the bench's 18 944 samples, 83 GB of outputs, one 20-warp group per SM -- skipped when the HBM is
    not free).  A random subset of sample columns is checked bit-for-bit against the oracle's propagation of the
    device's own durations and against the pair kernel run on the same seeds."""
    import torch

    gen = {"c2": synth.c2_layered, "c3": synth.c3_network}[name]
    dag, dists = gen()
    plan = capi.Plan(dag, dists, device=0)
    E, A = plan.E, plan.A
    free_b, _ = torch.cuda.mem_get_info()
    if (12 * E + 8 * A) * n > 0.8 * free_b:
        pytest.skip("not enough free HBM for the full-size launch")
    shape = plan.launch_shape(n)
    assert shape["samples_per_lane"] == 4, shape  # the auto rule picks the quad kernel for a machine-filling launch
    dev = torch.device("cuda:0")
    r = torch.empty((E, n), dtype=torch.float64, device=dev)
    d = torch.empty((A, n), dtype=torch.float64, device=dev)
    c = torch.empty((E, n), dtype=torch.int32, device=dev)
    seed0 = 123456
    plan.run_full_device(n, r, d, c, n, seed0=seed0)
    torch.cuda.synchronize()
    rng = np.random.default_rng(7)
    cols = np.sort(rng.choice(n, size=96, replace=False))
    cols[:4] = [0, 1, n - 2, n - 1]
    cols = np.unique(cols)
    idx = torch.as_tensor(cols, device=dev)
    r_s = r[:, idx].T.contiguous().cpu().numpy()
    d_s = d[:, idx].T.contiguous().cpu().numpy()
    c_s = c[:, idx].T.contiguous().cpu().numpy()
    del r, d, c
    torch.cuda.empty_cache()
    osim = oracle.OracleSim(dag, dists)
    r_o, c_o = osim.run_injected(d_s)
    assert np.array_equal(_bits(r_o), _bits(r_s)) and np.array_equal(c_o, c_s)
    r_p, d_p, c_p = _plan(dag, dists, 2).run_many_host((seed0 + cols).astype(np.int32))
    assert np.array_equal(_bits(d_p), _bits(d_s)) and np.array_equal(_bits(r_p), _bits(r_s)) and np.array_equal(c_p, c_s)


# ---- thread-block clusters: the CTAs of a cluster share one sample group and split its levels ------------------------

@pytest.mark.parametrize("cluster", [2, 4, 8])
@pytest.mark.parametrize("wpg", [2, 5, 20])
def test_cluster_launch_is_bit_identical(cluster, wpg):
    """Same bits with and without the cluster split, for injection, fused sampling (wide levels, long events that
    run through continuation chunks, orphan activities) and ragged sample counts."""
    dag = synth.random_dag(1200, 5, max_fan_in=40)
    dists = _all_kinds_dists()
    osim = oracle.OracleSim(dag, dists)
    seeds = np.arange(-5, 300, dtype=np.int32)  # 305 samples: three groups, the last one ragged
    _, dur, _ = osim.run_many(seeds)
    r_o, c_o = osim.run_injected(dur)
    ref = _plan(dag, dists, 4, wpg=wpg, gpc=1)
    ref.set_option(capi.OPT_CLUSTER_SIZE, 1)
    plan = _plan(dag, dists, 4, wpg=wpg, gpc=1)
    plan.set_option(capi.OPT_CLUSTER_SIZE, cluster)
    sh = plan.launch_shape(seeds.size)
    assert sh["cluster"] == cluster and sh["grid"] == 3 * cluster
    r_d, c_d = plan.run_injected_host(dur)
    assert np.array_equal(_bits(r_o), _bits(r_d)) and np.array_equal(c_o, c_d)
    a, b = ref.run_many_host(seeds), plan.run_many_host(seeds)
    for x, y in zip(a, b):
        assert np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8))


@pytest.mark.parametrize("cluster", [2, 8])
def test_cluster_launch_reduced_statistics(cluster):
    dag = synth.random_dag(900, 6, max_delay=60.0, max_fan_in=12)
    dists = synth.mixed_small_dists()
    seeds = np.arange(3, 3 + 700, dtype=np.int32)
    th = (1.0, 10.0)
    ref = _plan(dag, dists, 4, wpg=4, gpc=1)
    ref.set_option(capi.OPT_CLUSTER_SIZE, 1)
    plan = _plan(dag, dists, 4, wpg=4, gpc=1)
    plan.set_option(capi.OPT_CLUSTER_SIZE, cluster)
    a = ref.run_reduced_host(seeds, thresholds=th, n_bins=20, hist_range=(0.0, 60.0))
    b, act, none = plan.run_attribution_host(seeds, thresholds=th, n_bins=20, hist_range=(0.0, 60.0))
    np.testing.assert_allclose(b.sum, a.sum, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(b.sumsq, a.sumsq, rtol=1e-12, atol=1e-9)
    assert np.array_equal(a.late, b.late) and np.array_equal(a.hist, b.hist)
    _, act1, none1 = ref.run_attribution_host(seeds, thresholds=th, n_bins=20, hist_range=(0.0, 60.0))
    assert np.array_equal(act, act1) and np.array_equal(none, none1)


def test_auto_rule_clusters_small_launches_only():
    dag, dists = synth.c3_network()
    plan = capi.Plan(dag, dists, device=0)
    small, full = plan.launch_shape(2048), plan.launch_shape(18944)
    assert small["samples_per_lane"] == 4 and small["cluster"] == 4 and small["grid"] == 16 * 4
    assert full["cluster"] == 1 and full["grid"] == 148
    assert plan.launch_shape(1)["samples_per_lane"] == 1  # a call of one sample: the small-call path (test_gpu_small_calls.py)
    plan.set_option(capi.OPT_SMALL_CALL_MAX, 0)
    one = plan.launch_shape(1)
    assert one["samples_per_lane"] == 4 and one["cluster"] == 8 and one["grid"] == 8


@pytest.mark.parametrize("n_bins", [1, 7, 64, 65, 100])
def test_quad_histogram_staged_and_direct_paths(n_bins):
    """Up to 64 bins the warp stages 16-bit pair counts in shared memory, beyond that it aggregates with match.any:
    both equal the histogram of the full outputs (ragged sample count, event count not a multiple of the staging
    depth)."""
    dag = synth.random_dag(203, 9, max_delay=60.0)
    dists = synth.mixed_small_dists()
    seeds = np.arange(50, 50 + 333, dtype=np.int32)
    plan = _plan(dag, dists, 4)
    st = plan.run_reduced_host(seeds, thresholds=(0.0, 2.5, 59.0, 1e9), n_bins=n_bins, hist_range=(0.0, 60.0))
    r, _, _ = plan.run_many_host(seeds)
    delay = r - np.asarray(dag.earliest)[None, :]
    bins = np.clip(np.floor((delay - 0.0) * (n_bins / 60.0)).astype(np.int64), 0, n_bins - 1)
    ref = np.stack([np.bincount(bins[:, e], minlength=n_bins) for e in range(plan.E)])
    assert np.array_equal(st.hist.astype(np.int64), ref)
    for i, t in enumerate((0.0, 2.5, 59.0, 1e9)):
        assert np.array_equal(st.late[i], (delay > t).sum(0).astype(np.uint64))
    np.testing.assert_allclose(st.sum, delay.sum(0), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(st.sumsq, (delay * delay).sum(0), rtol=1e-12, atol=1e-9)
