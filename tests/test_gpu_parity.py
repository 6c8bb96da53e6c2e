"""GPU parity tests (run on the B200 with ``-m gpu``): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs.

Tier 1  duration injection: realized bit-exact in fp64, cause_event exact.
Tier 1b fused sampling kernel: its own durations reproduce its realized/cause through the
        oracle's propagation bit-for-bit; table-driven and constant draws equal the oracle's
        restatement of the device generator contract exactly, exponential/gamma to 1e-9 relative.
"""
import numpy as np
import pytest

import oracle
from mc_dagprop_b200 import capi, synth
from mc_dagprop_b200.flat import FlatDag, FlatDists

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _inject_case(dag, dists, seeds, wpg=0, gpc=0, chunk=0):
    osim = oracle.OracleSim(dag, dists)
    plan = capi.Plan(dag, dists, device=0)
    if wpg:
        plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
    if gpc:
        plan.set_option(capi.OPT_GROUPS_PER_CTA, gpc)
    if chunk:
        plan.set_option(capi.OPT_HOST_CHUNK, chunk)
    _, dur, _ = osim.run_many(seeds)  # reference-stream durations (Xoshiro256++)
    r_o, c_o = osim.run_injected(dur)
    r_d, c_d = plan.run_injected_host(dur)
    assert np.array_equal(_bits(r_o), _bits(r_d))
    assert np.array_equal(c_o, c_d)
    return plan, osim


@pytest.mark.parametrize("n_events,seed", [(2, 0), (17, 1), (200, 2), (1500, 3)])
@pytest.mark.parametrize("wpg", [1, 3, 8])
def test_injected_bit_exact_random_dags(n_events, seed, wpg):
    dag = synth.random_dag(n_events, seed)
    _inject_case(dag, synth.mixed_small_dists(), np.arange(-7, 150, dtype=np.int32), wpg=wpg)


def test_injected_reference_fixture():
    """reference test/test_simulator.py:107-137 golden vector, through the device."""
    dag = FlatDag.from_precedence_list(
        [0, 5, 10, 22, 20, 100], [(0, 3.0, 1), (1, 5.0, 1), (2, 5.0, 1), (3, 15.0, 2), (4, 10.0, 3)],
        [(1, [(0, 0)]), (2, [(1, 1)]), (3, [(1, 2)]), (4, [(2, 3), (3, 4)])], 1e6)
    d = FlatDists()
    d.add_constant(1, 1.0)
    d.add_constant(3, 3.0)
    plan = capi.Plan(dag, d, device=0)
    r, dur, c = plan.run_many_host([0, 1, 2, 3, 4])
    assert dur[3].tolist() == [6.0, 10.0, 10.0, 15.0, 40.0]
    assert r[3].tolist() == [0.0, 6.0, 16.0, 22.0, 62.0, 100.0]
    assert c[3].tolist() == [-1, 0, 1, -1, 3, -1]


def test_injected_special_values():
    """NaN / inf / signed zero durations follow std::min and >= exactly as the reference build."""
    dag = synth.random_dag(120, 5)
    dists = FlatDists()
    osim = oracle.OracleSim(dag, dists)
    plan = capi.Plan(dag, dists, device=0)
    rng = np.random.default_rng(0)
    dur = rng.integers(0, 20, size=(64, osim.A)).astype(np.float64)
    special = np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, 1e308, -1e308])
    mask = rng.random(dur.shape) < 0.1
    dur[mask] = rng.choice(special, size=int(mask.sum()))
    r_o, c_o = osim.run_injected(dur)
    r_d, c_d = plan.run_injected_host(dur)
    assert np.array_equal(_bits(r_o), _bits(r_d))
    assert np.array_equal(c_o, c_d)


@pytest.mark.parametrize("max_delay", [0.0, 3.0, float("inf")])
def test_injected_max_delay_edges(max_delay):
    dag = synth.random_dag(300, 7, max_delay=max_delay)
    _inject_case(dag, synth.mixed_small_dists(), np.arange(100, dtype=np.int32), wpg=2)


def test_injected_chunked_and_ragged():
    """n not a multiple of 64, several host chunks, odd chunk tails."""
    dag = synth.random_dag(150, 9)
    _inject_case(dag, synth.mixed_small_dists(), np.arange(1, 1000, dtype=np.int32), chunk=192)
    _inject_case(dag, synth.mixed_small_dists(), np.arange(1, dtype=np.int32))
    _inject_case(dag, synth.mixed_small_dists(), np.arange(63, dtype=np.int32))


def test_empty_inputs():
    dag = FlatDag.from_precedence_list([1.0, 2.0], [], [], 10.0)
    plan = capi.Plan(dag, FlatDists(), device=0)
    r, d, c = plan.run_many_host([1, 2, 3])
    assert d.shape == (3, 0) and r.tolist() == [[1.0, 2.0]] * 3 and c.tolist() == [[-1, -1]] * 3
    r, d, c = plan.run_many_host([])
    assert r.shape == (0, 2)


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("wpg", [1, 4])
def test_fused_kernel_self_consistent_and_matches_contract(seed, wpg):
    dag = synth.random_dag(400, 20 + seed)
    dists = synth.mixed_small_dists()
    osim = oracle.OracleSim(dag, dists)
    plan = capi.Plan(dag, dists, device=0)
    plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
    rng = np.random.default_rng(seed)
    # consecutive seeds (paired Philox blocks), odd start (unpaired), arbitrary, negative, extremes
    seeds = np.concatenate([np.arange(0, 130), np.arange(1001, 1100), rng.integers(-2**31, 2**31 - 1, size=100),
                            [-1, 2**31 - 1, -2**31]]).astype(np.int32)
    r, d, c = plan.run_many_host(seeds)
    r2, c2 = osim.run_injected(d)
    assert np.array_equal(_bits(r2), _bits(r))
    assert np.array_equal(c2, c)
    _, d_spec, _ = osim.run_many_spec(seeds)
    # activities whose draw involves no transcendental must agree bit-for-bit
    kinds = {t: k for t, k in zip(dists.dist_type.tolist(), dists.kind.tolist())}
    act_kind = np.full(osim.A, -1)
    for i, t in zip(dag.act_idx.tolist(), dag.act_type.tolist()):
        act_kind[i] = kinds.get(t, -1)
    exact = np.isin(act_kind, [-1, 0, 3, 4])
    assert np.array_equal(_bits(d[:, exact]), _bits(d_spec[:, exact]))
    # exponential: fp64 throughout (custom log vs libm): 1e-9 relative
    expo = act_kind == 1
    np.testing.assert_allclose(d[:, expo], d_spec[:, expo], rtol=1e-9, atol=1e-12)
    # gamma: the normal deviate and the accept tests use fp32 hardware approximations on the device, so
    # values agree to ~1e-5 relative and a borderline accept/reject may flip once in ~1e5 draws
    gam = act_kind == 2
    close = np.isclose(d[:, gam], d_spec[:, gam], rtol=5e-5, atol=1e-9)
    assert close.mean() > 0.9995, close.mean()
    # a sample is a pure function of its seed: permuting / re-partitioning the seeds changes nothing
    perm = rng.permutation(seeds.size)
    r3, d3, c3 = plan.run_many_host(seeds[perm])
    assert np.array_equal(_bits(d3), _bits(d[perm])) and np.array_equal(_bits(r3), _bits(r[perm]))
    assert np.array_equal(c3, c[perm])


def test_run_many_equals_single_runs():
    """reference test_monte_carlo_extra.py:94-109"""
    dag = synth.random_dag(60, 33)
    plan = capi.Plan(dag, synth.mixed_small_dists(), device=0)
    seeds = np.arange(5, dtype=np.int32)
    r, d, c = plan.run_many_host(seeds)
    for i, s in enumerate(seeds):
        r1, d1, c1 = plan.run_many_host([s])
        assert np.array_equal(_bits(r1[0]), _bits(r[i])) and np.array_equal(_bits(d1[0]), _bits(d[i]))
        assert np.array_equal(c1[0], c[i])


def test_reduced_mode_matches_full_outputs():
    dag = synth.random_dag(500, 41, max_delay=40.0)
    dists = synth.mixed_small_dists()
    plan = capi.Plan(dag, dists, device=0)
    seeds = np.arange(3, 3 + 2000, dtype=np.int32)
    r, _, _ = plan.run_many_host(seeds)
    delay = r - dag.earliest[None, :]
    th = (1.0, 5.0, 20.0)
    for wpg in (1, 4):
        plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
        st = plan.run_reduced_host(seeds, thresholds=th, n_bins=16, hist_range=(0.0, 40.0))
        np.testing.assert_allclose(st.sum, delay.sum(0), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(st.sumsq, (delay * delay).sum(0), rtol=1e-12, atol=1e-9)
        for i, t in enumerate(th):
            assert np.array_equal(st.late[i], (delay > t).sum(0).astype(np.uint64))
        bins = np.clip(np.floor(delay * (16 / 40.0)).astype(int), 0, 15)
        ref_hist = np.stack([np.bincount(bins[:, e], minlength=16) for e in range(plan.E)])
        assert np.array_equal(st.hist, ref_hist.astype(np.uint32))
        assert int(st.hist.sum()) == seeds.size * plan.E


@pytest.mark.parametrize("seed0", [0, 4096, 7, -6])
@pytest.mark.parametrize("wpg", [0, 1, 4])
def test_device_call_with_seed0_equals_seed_array(seed0, wpg):
    """mcdp_run_full_device / mcdp_run_reduced_device with seeds given as seed0 + column (the bench path; lanes own
    seed pairs {2k, 2k+1} when seed0 is even) return exactly what the same seeds give as an explicit array, in an
    order in which no lane owns a pair."""
    import torch

    dag, dists = synth.random_dag(400, 21), synth.mixed_small_dists()
    plan = capi.Plan(dag, dists, device=0)
    if wpg:
        plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
    n, ld = 300, 320
    E, A = plan.E, plan.A
    dev = torch.device("cuda:0")
    r = torch.zeros((E, ld), dtype=torch.float64, device=dev)
    d = torch.zeros((A, ld), dtype=torch.float64, device=dev)
    c = torch.zeros((E, ld), dtype=torch.int32, device=dev)
    plan.run_full_device(n, r, d, c, ld, seed0=seed0)
    torch.cuda.synchronize()
    seeds = np.arange(seed0, seed0 + n, dtype=np.int32)
    # reversed order: no lane owns a seed pair (one Philox block per sample instead of one per pair)
    r_h, d_h, c_h = (x[::-1] for x in plan.run_many_host(seeds[::-1].copy()))
    r_p, d_p, c_p = plan.run_many_host(seeds)  # in order: pairs wherever seed0 is even
    assert np.array_equal(_bits(r_p), _bits(r_h)) and np.array_equal(_bits(d_p), _bits(d_h)) and np.array_equal(c_p, c_h)
    assert np.array_equal(_bits(r[:, :n].T.cpu().numpy()), _bits(r_h))
    assert np.array_equal(_bits(d[:, :n].T.cpu().numpy()), _bits(d_h))
    assert np.array_equal(c[:, :n].T.cpu().numpy(), c_h)
    # reduced statistics: integer accumulators exact, sums to summation order
    desc = capi.make_stats_desc(thresholds=(1.0, 10.0), n_bins=16, hist_range=(0.0, 60.0))
    s_sum = torch.zeros(E, dtype=torch.float64, device=dev)
    s_sq = torch.zeros(E, dtype=torch.float64, device=dev)
    s_late = torch.zeros((2, E), dtype=torch.int64, device=dev)
    s_hist = torch.zeros((E, 16), dtype=torch.int32, device=dev)
    plan.run_reduced_device(n, desc, s_sum, s_sq, s_late, s_hist, seed0=seed0)
    torch.cuda.synchronize()
    st = plan.run_reduced_host(seeds, thresholds=(1.0, 10.0), n_bins=16, hist_range=(0.0, 60.0))
    assert np.array_equal(s_late.cpu().numpy().astype(np.uint64), np.asarray(st.late, dtype=np.uint64))
    assert np.array_equal(s_hist.cpu().numpy().astype(np.uint32), np.asarray(st.hist, dtype=np.uint32))
    assert np.allclose(s_sum.cpu().numpy(), st.sum, rtol=1e-12, atol=1e-9)
    assert np.allclose(s_sq.cpu().numpy(), st.sumsq, rtol=1e-12, atol=1e-9)
