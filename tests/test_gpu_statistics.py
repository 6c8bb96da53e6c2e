"""Tier 2 and tier 3 parity (``-m gpu``).

Tier 2: per-distribution samples of the device generator (Philox contract) vs the reference's
        generator (Xoshiro256++ + libstdc++ transforms, through the pinned oracle): two-sample
        Kolmogorov-Smirnov, p > 1e-3 at ~130k draws per arm (stated tolerance).
Tier 3: per-event realized-time distributions on the same DAG: quantiles q in {0.1 .. 0.99} of the
        device run vs the reference stream agree within 3 standard errors of the quantile estimate
        plus 1e-9 (stated tolerance), and means within 4 standard errors.
"""
import numpy as np
import pytest
from scipy import stats

import oracle
from mc_dagprop_b200 import capi, synth
from mc_dagprop_b200.flat import FlatDag, FlatDists

pytestmark = pytest.mark.gpu


def _star(n_links, base, add):
    """1 source, n parallel links of one type into n sinks (pattern of reference demo/distribution.py:17-36)."""
    earliest = np.zeros(n_links + 1)
    acts = [(i, base, 1) for i in range(n_links)]
    prec = [(i + 1, [(0, i)]) for i in range(n_links)]
    d = FlatDists()
    add(d)
    return FlatDag.from_precedence_list(earliest, acts, prec, 1e9), d


CASES = {
    "exponential": lambda d: d.add_exponential(1, 0.7, 50.0),
    "exponential_truncated": lambda d: d.add_exponential(1, 2.0, 1.5),
    "exponential_heavily_truncated": lambda d: d.add_exponential(1, 100.0, 1.0),
    "gamma_shape_2": lambda d: d.add_gamma(1, 2.0, 0.1, 5.0),
    "gamma_shape_half": lambda d: d.add_gamma(1, 0.5, 0.3, 5.0),
    "gamma_shape_7_untruncated": lambda d: d.add_gamma(1, 7.3, 0.2),
    "gamma_truncated": lambda d: d.add_gamma(1, 2.0, 1.0, 1.5),
    # 2 * shape in {1..6, 8} takes the exact-transformation path (sum of exponentials + half a squared normal) ...
    "gamma_shape_1": lambda d: d.add_gamma(1, 1.0, 0.4, 6.0),
    "gamma_shape_1p5": lambda d: d.add_gamma(1, 1.5, 0.4, 6.0),
    "gamma_shape_2p5_truncated": lambda d: d.add_gamma(1, 2.5, 0.5, 1.5),
    "gamma_shape_3p5_truncated": lambda d: d.add_gamma(1, 3.5, 0.5, 2.0),  # five uniforms: Marsaglia-Tsang
    "gamma_shape_4": lambda d: d.add_gamma(1, 4.0, 0.25),
    # ... every other shape Marsaglia-Tsang
    "gamma_shape_2p3": lambda d: d.add_gamma(1, 2.3, 0.1, 5.0),
    "gamma_shape_0p6_truncated": lambda d: d.add_gamma(1, 0.6, 0.5, 1.0),
    "empirical_relative_256": lambda d: d.add_empirical_relative(1, np.linspace(0, 3, 256), np.exp(-np.linspace(0, 3, 256))),
    "empirical_absolute_5": lambda d: d.add_empirical_absolute(1, [0.0, 1.0, 2.0, 5.0, 9.0], [0.3, 0.3, 0.2, 0.15, 0.05]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_tier2_ks_against_reference_generator(name):
    dag, d = _star(64, 3.0, CASES[name])
    seeds = np.arange(2048, dtype=np.int32)
    _, dur_ref, _ = oracle.OracleSim(dag, d).run_many(seeds)
    _, dur_dev, _ = capi.Plan(dag, d, device=0).run_many_host(seeds)
    a, b = (dur_ref - 3.0).ravel(), (dur_dev - 3.0).ravel()
    if name.startswith("empirical"):
        va, ca = np.unique(a, return_counts=True)
        vb, cb = np.unique(b, return_counts=True)
        assert set(vb.tolist()) <= set(va.tolist()) | set(np.round(vb, 12).tolist())
        allv = np.union1d(va, vb)
        fa = np.array([ca[va == v].sum() for v in allv]) / a.size
        fb = np.array([cb[vb == v].sum() for v in allv]) / b.size
        assert np.abs(np.cumsum(fa) - np.cumsum(fb)).max() < 0.01
    else:
        res = stats.ks_2samp(a, b)
        assert res.pvalue > 1e-3, (name, res)
        assert abs(a.mean() - b.mean()) < 5 * a.std() / np.sqrt(a.size) * np.sqrt(2)


def test_tier2_each_activity_has_its_own_stream():
    """Draws of different activities / seeds are uncorrelated (counter-based keying works)."""
    dag, d = _star(32, 1.0, lambda g: g.add_exponential(1, 1.0, 1e9))
    _, dur, _ = capi.Plan(dag, d, device=0).run_many_host(np.arange(4096, dtype=np.int32))
    x = dur - 1.0
    c = np.corrcoef(x.T)
    off = c[~np.eye(32, dtype=bool)]
    assert np.abs(off).max() < 0.08
    c2 = np.corrcoef(x[0::2, 0], x[1::2, 0])[0, 1]  # the two halves of a PAIR block
    assert abs(c2) < 0.06


@pytest.mark.parametrize("gen", ["c2", "c3", "random"])
def test_tier3_per_event_quantiles(gen):
    if gen == "c2":
        dag, d = synth.c2_layered(12, 16)
    elif gen == "c3":
        dag, d = synth.c3_network(10, 24)
    else:
        dag, d = synth.random_dag(150, 77, tie_prone=False, max_delay=200.0), synth.mixed_small_dists()
    n = 20_000
    seeds = np.arange(n, dtype=np.int32)
    r_ref, _, _ = oracle.OracleSim(dag, d).run_many(seeds, durations=False, cause=False)
    r_dev, _, _ = capi.Plan(dag, d, device=0).run_many_host(seeds + 1_000_000, durations=False, cause=False)
    # means: 4 standard errors of the difference
    se = np.sqrt(r_ref.var(0) / n + r_dev.var(0) / n)
    assert np.all(np.abs(r_ref.mean(0) - r_dev.mean(0)) <= 4 * se + 1e-9)
    # quantiles: compare via the CDF -- the fraction of device samples below the reference quantile
    # must be q within 4 binomial standard errors (ties make value-space comparison ill-posed)
    worst = 0.0
    for q in (0.1, 0.25, 0.5, 0.75, 0.9, 0.99):
        xq = np.quantile(r_ref, q, axis=0)
        lo = (r_dev < xq[None, :] - 1e-9).mean(0)
        hi = (r_dev <= xq[None, :] + 1e-9).mean(0)
        lo_ref = (r_ref < xq[None, :] - 1e-9).mean(0)
        hi_ref = (r_ref <= xq[None, :] + 1e-9).mean(0)
        tol = 4 * np.sqrt(2 * q * (1 - q) / n) + 2.0 / n
        bad = (lo > hi_ref + tol) | (hi < lo_ref - tol)
        worst = max(worst, float(np.max(np.maximum(lo - hi_ref, lo_ref - hi))))
        assert not bad.any(), (gen, q, int(bad.sum()))
    assert worst < 0.05


def _expected_cause_counts(dag, realized, durations):
    """Replays _core.cpp:332-350 on given realized / durations samples and counts, per activity, how often its
    precedence entry decided the target event (the last entry whose clamped arrival reached the running maximum)."""
    n, E = realized.shape
    A = durations.shape[1]
    cause_act = np.zeros(A, np.uint64)
    cause_none = np.zeros(E, np.uint64)
    entry_of = {}
    for i, t in enumerate(dag.prec_target):  # the last entry for a target wins (_core.cpp:240)
        entry_of[int(t)] = i
    for e in range(E):
        ub = dag.earliest[e] + dag.max_delay
        latest = np.full(n, dag.earliest[e])
        win = np.full(n, -1, np.int64)
        i = entry_of.get(e)
        if i is not None:
            for k in range(dag.prec_off[i], dag.prec_off[i + 1]):
                src, act = int(dag.pred_src[k]), int(dag.pred_act[k])
                dur = durations[:, act] if act < A else 0.0
                t = np.minimum(realized[:, src] + dur, ub)
                take = t >= latest
                latest = np.where(take, t, latest)
                win = np.where(take, act if act < A else -3, win)
        assert np.array_equal(np.minimum(latest, ub).view(np.uint64), realized[:, e].copy().view(np.uint64))
        cause_none[e] = np.count_nonzero(win == -1)
        acts, counts = np.unique(win[win >= 0], return_counts=True)
        cause_act[acts] += counts.astype(np.uint64)
    return cause_act, cause_none


@pytest.mark.parametrize("wpg", [0, 1])
def test_delay_cause_attribution_counts(wpg):
    """mcdp_run_attribution_host: per-activity counts of being the binding predecessor equal the counts derived
    from the full outputs of the same seeds, exactly; the other statistics are those of the plain reduced call."""
    dag, dists = synth.random_dag(250, 5, max_delay=40.0), synth.mixed_small_dists()
    plan = capi.Plan(dag, dists, device=0)
    if wpg:
        plan.set_option(capi.OPT_WARPS_PER_GROUP, wpg)
    seeds = np.arange(3, 3 + 1000, dtype=np.int32)
    r, d, c = plan.run_many_host(seeds)
    exp_act, exp_none = _expected_cause_counts(dag, r, d)
    assert np.array_equal(exp_none, (c == -1).sum(0).astype(np.uint64))
    st, cause_act, cause_none = plan.run_attribution_host(seeds, thresholds=(1.0,), n_bins=8, hist_range=(0.0, 40.0))
    assert np.array_equal(cause_act, exp_act)
    assert np.array_equal(cause_none, exp_none)
    assert int(cause_act.sum() + cause_none.sum()) <= seeds.size * plan.E
    ref = plan.run_reduced_host(seeds, thresholds=(1.0,), n_bins=8, hist_range=(0.0, 40.0))
    assert np.array_equal(st.late, ref.late) and np.array_equal(st.hist, ref.hist)
    assert np.allclose(st.sum, ref.sum, rtol=1e-12, atol=1e-9)
