"""Pins the CPU oracle (oracle/mcdp_oracle.c) to the reference: (1) the golden vectors the
reference's own tests hold, (2) committed fixtures generated from the unmodified reference
(tests/golden/make_golden.py), (3) live bit-for-bit comparison with oracle/_ref when it is
present.  CPU only."""
import os

import numpy as np
import pytest

import oracle
from mc_dagprop_b200 import synth
from mc_dagprop_b200.flat import FlatDag, FlatDists

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _fixture_dag(max_delay=1e6):
    """reference test/test_simulator.py:11-35"""
    return FlatDag.from_precedence_list(
        [0.0, 5.0, 10.0, 22.0, 20.0, 100.0],
        [(0, 3.0, 1), (1, 5.0, 1), (2, 5.0, 1), (3, 15.0, 2), (4, 10.0, 3)],
        [(1, [(0, 0)]), (2, [(1, 1)]), (3, [(1, 2)]), (4, [(2, 3), (3, 4)])], max_delay)


def test_golden_propagation_vector():
    """reference test/test_simulator.py:107-137"""
    d = FlatDists()
    d.add_constant(1, 1.0)
    d.add_constant(3, 3.0)
    r, dur, c = oracle.OracleSim(_fixture_dag(), d).run_many(range(5))
    assert dur[3].tolist() == [6.0, 10.0, 10.0, 15.0, 40.0]
    assert r[3].tolist() == [0.0, 6.0, 16.0, 22.0, 62.0, 100.0]
    assert c[3].tolist() == [-1, 0, 1, -1, 3, -1]


def test_golden_empirical_known_answers():
    """reference test/test_simulator.py:139-155: Xoshiro256++ + generate_canonical + discrete_distribution"""
    d = FlatDists()
    d.add_empirical_absolute(1, [10, 20, 40, 50], [0.1, 0.2, 0.3, 0.4])
    r, _, _ = oracle.OracleSim(_fixture_dag(), d).run_many([7])
    assert r[0, 3] == 68.0 and r[0, 5] == 100.0
    d = FlatDists()
    d.add_empirical_relative(1, [1.2, 1.3, 1.35, 4.5], [0.1, 0.2, 0.3, 0.4])
    r, _, _ = oracle.OracleSim(_fixture_dag(), d).run_many([7])
    assert abs(r[0, 3] - 34.10) < 5e-5 and r[0, 5] == 100.0


def test_golden_empirical_relative_with_exponential_histogram():
    """reference test/test_simulator.py:157-172 (numpy legacy stream + 1000-bin histogram)"""
    # the reference draws one value at a time and keeps those <= 5.0; block draws give the same stream
    np.random.seed(7)
    out = []
    need = 1_000_000
    while len(out) < need:
        block = np.random.exponential(3.0, size=need - len(out))
        out.extend(block[block <= 5.0].tolist())
    values = np.array(out[:need])
    hist, edges = np.histogram(values, bins=1000, density=True)
    centers = 0.5 * (edges[:-1] + edges[1:])
    d = FlatDists()
    d.add_empirical_relative(1, centers, hist)
    r, _, _ = oracle.OracleSim(_fixture_dag(), d).run_many([7])
    assert abs(r[0, 3] - 23.062461393412335) < 5e-4 and r[0, 5] == 100.0


def test_golden_constant_chain_and_large_scale():
    """reference test/test_monte_carlo_extra.py:27-34 and test/test_simulator.py:184-199"""
    dag = FlatDag.from_precedence_list([0.0, 0.0, 0.0], [(0, 1.0, 1), (1, 2.0, 1)], [(1, [(0, 0)]), (2, [(1, 1)])], 1e6)
    d = FlatDists()
    d.add_constant(1, 1.0)
    r, dur, _ = oracle.OracleSim(dag, d).run_many([42])
    assert dur[0].tolist() == [2.0, 4.0] and r[0].tolist() == [0.0, 2.0, 6.0]
    n = 10_000
    dag = FlatDag.from_precedence_list(np.arange(n, dtype=float), [(i, 3.0, 1) for i in range(n - 1)],
                                       [(i, [(i - 1, i)]) for i in range(1, n)], 1e6)
    r, dur, c = oracle.OracleSim(dag, d).run_many([7])
    assert r.shape == (1, n) and dur.shape == (1, n - 1) and r[0, 0] == 0.0 and r[0, n - 1] == 59988.0


def test_reference_error_conditions():
    """reference test/test_monte_carlo_extra.py:60-91"""
    d = FlatDists()
    d.add_constant(1, 0.0)
    cyc = FlatDag.from_precedence_list([0.0, 0.0], [(0, 1.0, 1), (1, 1.0, 1)], [(1, [(0, 0)]), (0, [(1, 1)])], 1e6)
    with pytest.raises(RuntimeError, match="cycle"):
        oracle.OracleSim(cyc, d)
    ok = FlatDag.from_precedence_list([0.0, 0.0], [(0, 1.0, 1)], [(1, [(0, 0)])], -1.0)
    with pytest.raises(RuntimeError, match="max_delay"):
        oracle.OracleSim(ok, d)
    d2 = FlatDists()
    d2.add_constant(-1, 0.0)
    ok.max_delay = 1.0
    with pytest.raises(RuntimeError, match="reserved"):
        oracle.OracleSim(ok, d2)


def test_probed_semantics_table():
    """SURVEY.md section 8(a) 'Semantics that the build must reproduce exactly'."""
    none = FlatDists()

    def run(earliest, acts, prec, md):
        r, d, c = oracle.OracleSim(FlatDag.from_precedence_list(earliest, acts, prec, md), none).run_many([0])
        return r[0].tolist(), d[0].tolist(), c[0].tolist()

    # pred arrives exactly at earliest -> cause = src (>=)
    r, _, c = run([0.0, 5.0], [(0, 5.0, 0)], [(1, [(0, 0)])], 100.0)
    assert r == [0.0, 5.0] and c == [-1, 0]
    # two preds tie -> last listed wins
    r, _, c = run([0.0, 0.0, 1.0], [(0, 3.0, 0), (1, 3.0, 0)], [(2, [(0, 0), (1, 1)])], 100.0)
    assert c[2] == 1
    r, _, c = run([0.0, 0.0, 1.0], [(0, 3.0, 0), (1, 3.0, 0)], [(2, [(1, 1), (0, 0)])], 100.0)
    assert c[2] == 0
    # several preds exceed ub -> clamped before comparison, last listed that reaches ub wins
    r, _, c = run([0.0, 0.0, 10.0], [(0, 500.0, 0), (1, 900.0, 0)], [(2, [(1, 1), (0, 0)])], 100.0)
    assert r[2] == 110.0 and c[2] == 0
    # max_delay = 0 -> realized = earliest everywhere
    r, _, c = run([0.0, 0.0, 10.0], [(0, 500.0, 0), (1, 2.0, 0)], [(2, [(0, 0), (1, 1)])], 0.0)
    assert r == [0.0, 0.0, 10.0] and c[2] == 0
    # idx gaps become zero-duration links
    r, d, _ = run([0.0, 1.0], [(3, 7.0, 0)], [(1, [(0, 3)])], 100.0)
    assert d == [0.0, 0.0, 0.0, 7.0] and r == [0.0, 7.0]
    # empty activities / empty seeds
    sim = oracle.OracleSim(FlatDag.from_precedence_list([1.0], [], [], 5.0), none)
    r, d, c = sim.run_many([])
    assert r.shape == (0, 1) and d.shape == (0, 0)


def test_rng_restatement_known_values():
    """Xoshiro256++ seeded through SplitMix64 (_custom_rng.hpp:507-600); canonical in [0,1)."""
    g = oracle.Xoshiro(0)
    a = [g.next() for _ in range(3)]
    # SplitMix64(0) first outputs are the published test vector of the algorithm
    sm = [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F, 0xF88BB8A8724C81EC]
    s = list(sm)
    exp = []
    for _ in range(3):
        rotl = lambda x, k: ((x << k) | (x >> (64 - k))) & (2**64 - 1)  # noqa: E731
        res = (rotl((s[0] + s[3]) & (2**64 - 1), 23) + s[0]) & (2**64 - 1)
        t = (s[1] << 17) & (2**64 - 1)
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45)  # noqa: E702
        exp.append(res)
    assert a == exp
    g = oracle.Xoshiro(-1 & (2**64 - 1))
    u = [g.canonical() for _ in range(1000)]
    assert 0.0 <= min(u) and max(u) < 1.0


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32-10 (device generator contract, not reference behaviour)."""
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_fixtures_random_dags_bit_exact():
    """Committed outputs of the unmodified reference for seeded random DAGs, every distribution kind,
    two consecutive run_many passes (gamma's cached normal survives reseeding: SURVEY 8a row D3)."""
    from tests.golden.make_golden import CASES, SEEDS

    fx = np.load(os.path.join(GOLD, "ref_random_dags.npz"))
    for n, s in CASES:
        for md in (50.0, 4.0):
            key = f"n{n}_s{s}_md{int(md)}"
            sim = oracle.OracleSim(synth.random_dag(n, s, max_delay=md), synth.mixed_small_dists())
            for suffix in ("", "2"):
                r, d, c = sim.run_many(SEEDS)
                assert np.array_equal(_bits(r), _bits(fx[key + "_realized" + suffix])), key
                assert np.array_equal(_bits(d), _bits(fx[key + "_durations" + suffix])), key
                assert np.array_equal(c, fx[key + "_cause" + suffix]), key


def test_fixtures_python_api_bit_exact():
    """Outputs recorded through the reference's own pybind11 API (test_simulator.py fixture)."""
    fx = np.load(os.path.join(GOLD, "ref_python_api.npz"))
    dag = _fixture_dag()

    def check(name, dists, seeds):
        r, d, c = oracle.OracleSim(dag, dists).run_many(seeds)
        assert np.array_equal(_bits(r), _bits(fx[name + "_realized"])), name
        assert np.array_equal(_bits(d), _bits(fx[name + "_durations"])), name
        assert np.array_equal(c, fx[name + "_cause"]), name

    d = FlatDists(); d.add_constant(1, 1.0); d.add_constant(3, 3.0)  # noqa: E702
    check("constant", d, range(5))
    d = FlatDists(); d.add_empirical_absolute(1, [10, 20, 40, 50], [0.1, 0.2, 0.3, 0.4])  # noqa: E702
    check("emp_abs", d, range(16))
    d = FlatDists(); d.add_empirical_relative(1, [1.2, 1.3, 1.35, 4.5], [0.1, 0.2, 0.3, 0.4])  # noqa: E702
    check("emp_rel", d, range(16))
    d = FlatDists(); d.add_exponential(1, 1000.0, 1.0)  # noqa: E702
    check("exp_rejection_heavy", d, range(3))
    d = FlatDists(); d.add_gamma(1, 2.0, 1.0, 0.5)  # noqa: E702
    check("gamma_truncated", d, range(8))


@pytest.mark.parametrize("n,seed", [(30, 1), (300, 2), (1200, 3)])
def test_live_against_reference_build(n, seed, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    dag, dists = synth.random_dag(n, seed), synth.mixed_small_dists()
    o, r = oracle.OracleSim(dag, dists), oracle.RefSim(dag, dists)
    seeds = np.arange(-50, 150, dtype=np.int32)
    for _ in range(2):
        a, b = o.run_many(seeds), r.run_many(seeds)
        assert np.array_equal(_bits(a[0]), _bits(b[0])) and np.array_equal(_bits(a[1]), _bits(b[1]))
        assert np.array_equal(a[2], b[2])


def test_live_synthetic_configs_against_reference_build(have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for gen in (synth.c1_toy, lambda: synth.c1_toy("const_exp"), lambda: synth.c2_layered(20, 30),
                lambda: synth.c3_network(12, 40), lambda: synth.c5_deep_chain(300, 4, 32)):
        dag, dists = gen()
        a, b = oracle.OracleSim(dag, dists).run_many(range(40)), oracle.RefSim(dag, dists).run_many(range(40))
        assert np.array_equal(_bits(a[0]), _bits(b[0])) and np.array_equal(_bits(a[1]), _bits(b[1]))
        assert np.array_equal(a[2], b[2])
