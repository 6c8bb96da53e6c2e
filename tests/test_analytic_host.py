"""Host side of the analytic package (no GPU): the discretised distributions against the reference fixture, the
reference's validation behaviour, and the package surface (reference ``analytic/__init__.py``)."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_analytic.npz"))


def test_distribution_builders_match_reference_fixture():
    from mc_dagprop.analytic import constant_pmf, empirical_pmf, exponential_pmf, gamma_pmf

    np.testing.assert_allclose(exponential_pmf(scale=10.0, step=1, start=0.0, stop=300.0).probabilities, GOLD["dist_exponential"], rtol=1e-13)
    np.testing.assert_allclose(gamma_pmf(shape=2.0, scale=2.0, step=1, start=0.0, stop=40.0).probabilities, GOLD["dist_gamma"], rtol=1e-11)
    np.testing.assert_allclose(gamma_pmf(shape=0.5, scale=3.0, step=2, start=0.0, stop=60.0).probabilities, GOLD["dist_gamma_half"], rtol=1e-11)
    np.testing.assert_allclose(empirical_pmf([0.0, 1.0, 2.0], [1, 1, 2], step=1).probabilities, GOLD["dist_empirical"], rtol=1e-15)
    c = constant_pmf(5.0, step=1)
    assert np.allclose(c.values, [5.0]) and np.allclose(c.probabilities, [1.0])
    for bad in (lambda: exponential_pmf(0.0, 1, 0, 10), lambda: exponential_pmf(1.0, 0, 0, 10), lambda: gamma_pmf(1.0, 1.0, 1, 5, 0),
                lambda: empirical_pmf([0.0], [0.0], 1), lambda: empirical_pmf([0.0, 1.0], [1.0], 1)):
        with pytest.raises(ValueError):
            bad()


def test_pmf_checks_follow_the_reference():
    from mc_dagprop import DiscretePMF

    with pytest.raises(ValueError, match="cannot be empty"):
        DiscretePMF(np.array([]), np.array([]), step=1)
    with pytest.raises(ValueError, match="same length"):
        DiscretePMF(np.array([0.0, 1.0]), np.array([1.0]), step=1)
    with pytest.raises(ValueError, match="sorted"):
        DiscretePMF(np.array([1.0, 0.0]), np.array([0.5, 0.5]), step=1)
    with pytest.raises(ValueError, match="sum to <= 1.0"):
        DiscretePMF(np.array([0.0, 1.0]), np.array([0.8, 0.8]), step=1)
    with pytest.raises(OverflowError):
        DiscretePMF(np.array([0.0]), np.array([1.0]), step=1.0)
    p = DiscretePMF(np.array([2.0, 3.0]), np.array([0.25, 0.5]), step=1)
    assert float(p.total_mass) == 0.75 and np.array_equal(p.shift(2).values, [4.0, 5.0])
    p.validate_alignment(1.0)
    with pytest.raises(ValueError, match="does not match expected"):
        p.validate_alignment(2.0)
    with pytest.raises(ValueError, match="grid spacing"):
        DiscretePMF(np.array([0.0, 2.0]), np.array([0.5, 0.5]), step=1).validate_alignment(1)
    d = DiscretePMF.delta(7.0, 1)
    assert d.values.tolist() == [7.0] and d.probabilities.tolist() == [1.0]


def test_context_validation_follows_the_reference():
    from mc_dagprop import AnalyticContext, DiscretePMF, Event, EventTimestamp, create_analytic_propagator
    from mc_dagprop.analytic import AnalyticActivity, OverflowRule, UnderflowRule

    ev = tuple(Event(str(i), EventTimestamp(0.0, 100.0, 0.0)) for i in range(3))
    act = AnalyticActivity(0, DiscretePMF(np.array([1.0, 2.0]), np.array([0.5, 0.5]), step=1))

    def ctx(**kw):
        base = dict(events=ev, activities={(0, 1): (0, act)}, precedence_list=((1, ((0, 0),)),), step=1,
                    underflow_rule=UnderflowRule.TRUNCATE, overflow_rule=OverflowRule.TRUNCATE)
        base.update(kw)
        return AnalyticContext(**base)

    sim = create_analytic_propagator(ctx())  # valid: builds without a GPU
    assert sim._topological_node_order == (0, 2, 1) and sim._predecessors_by_target[1] == ((0, 0),)
    assert sim._event_bounds(3.4, 50.0) == (3, 50)
    assert create_analytic_propagator(ctx(max_delay=10))._event_bounds(3.4, 50.0) == (3, 13)
    for bad in (ctx(step=0), ctx(step=2), ctx(max_delay=-1), ctx(precedence_list=((1, ((0, 5),)),)),
                ctx(precedence_list=((1, ((2, 0),)),)), ctx(precedence_list=((7, ()),)),
                ctx(events=(Event("x", EventTimestamp(5.0, 4.0, 0.0)),), activities={}, precedence_list=()),
                ctx(activities={(0, 1): (0, act), (1, 0): (1, act)}, precedence_list=((1, ((0, 0),)), (0, ((1, 1),)))),
                ctx(activities={(0, 1): (0, AnalyticActivity(0, DiscretePMF(np.array([1.0, 2.5]), np.array([0.5, 0.5]), step=1)))})):
        with pytest.raises(ValueError):
            create_analytic_propagator(bad)
    create_analytic_propagator(ctx(step=2), validate=False)  # validation can be skipped, as in the reference


def test_package_surface_matches_reference():
    import mc_dagprop
    from mc_dagprop import analytic

    for name in ("DiscretePMF", "SimulatedEvent", "UnderflowRule", "OverflowRule", "AnalyticContext", "AnalyticPropagator",
                 "create_analytic_propagator"):
        assert name in mc_dagprop.__all__ and getattr(mc_dagprop, name) is getattr(analytic, name)
    for name in ("AnalyticActivity", "exponential_pmf", "gamma_pmf", "constant_pmf", "empirical_pmf", "Second", "ProbabilityMass",
                 "EventIndex", "ActivityIndex"):
        assert name in analytic.__all__
    assert [int(r) for r in analytic.UnderflowRule] == [1, 2, 3] and analytic.OverflowRule.REDISTRIBUTE == 3
    from mc_dagprop.analytic._context import AnalyticActivity, SimulatedEvent  # noqa: F401  (paths the reference's tests import)
    from mc_dagprop.analytic._pmf import DiscretePMF  # noqa: F401


@pytest.mark.parametrize("slack", [True, False])
def test_run_unpacks_the_device_result(monkeypatch, slack):
    """``AnalyticPropagator.run`` around a stand-in for the device call: result slots longer than the results, value
    grids from start / step, the bulk mass check of the reference's ``DiscretePMF`` constructor."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import types

    import analytic_cases as ac
    import mc_dagprop
    import mc_dagprop.analytic as an
    from mc_dagprop_b200.analytic import _device

    ns = types.SimpleNamespace(Event=mc_dagprop.Event, EventTimestamp=mc_dagprop.EventTimestamp, DiscretePMF=an.DiscretePMF,
                               AnalyticActivity=an.AnalyticActivity, AnalyticContext=an.AnalyticContext,
                               UnderflowRule=an.UnderflowRule, OverflowRule=an.OverflowRule)
    ctx = ac.build_context(ns, "dag", 120, 3, 5, None, 1, 1)
    rng = np.random.default_rng(0)
    sent = {}

    def fake_run(lower, upper, origin, step, *rest):
        n = len(lower)
        length = rng.integers(1, 30, n).astype(np.int32)
        slots = length + (rng.integers(0, 4, n) if slack else 0)
        off = np.concatenate([[0], np.cumsum(slots)]).astype(np.int64)
        probs = rng.random(int(off[-1])) + 5.0  # whatever lies in the unused tail of a slot must not be looked at
        for i in range(n):
            seg = probs[off[i]: off[i] + length[i]]
            seg /= seg.sum() * (1.0 + 1e-9)
        under = rng.random(n) * 1e-3
        sent.update(start=np.asarray(lower), length=length, off=off, probs=probs.copy(), under=under, step=step)
        return np.asarray(lower), length, off, probs, under, np.zeros(n)

    monkeypatch.setattr(_device, "analytic_run", fake_run)
    prop = an.create_analytic_propagator(ctx)
    res = prop.run()
    assert len(res) == len(ctx.events) and sent["step"] == 5
    for i, r in enumerate(res):
        k, o = int(sent["length"][i]), int(sent["off"][i])
        assert np.array_equal(r.pmf.probabilities, sent["probs"][o:o + k])
        assert np.array_equal(r.pmf.values, float(sent["start"][i]) + 5.0 * np.arange(k))
        assert r.pmf.step == 5 and float(r.underflow) == sent["under"][i] and float(r.overflow) == 0.0
        r.pmf.validate()
    # a result whose mass exceeds one is refused like the reference's constructor refuses it
    def too_heavy(*args):
        out = list(fake_run(*args))
        out[3][int(out[2][7]): int(out[2][7]) + int(out[1][7])] *= 1.5
        return tuple(out)

    monkeypatch.setattr(_device, "analytic_run", too_heavy)
    with pytest.raises(ValueError, match="sum to <= 1.0"):
        prop.run()
