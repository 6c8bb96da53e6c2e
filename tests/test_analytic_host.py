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
