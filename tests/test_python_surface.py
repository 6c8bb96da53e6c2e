"""The drop-in Python surface (reference _core.pyi / _core.cpp:366-552), checked without a GPU."""
import dataclasses

import pytest

import mc_dagprop
from mc_dagprop import (Activity, DagContext, Event, EventTimestamp, GenericDelayGenerator, MonteCarloPropagator,
                        SimResult, Simulator)
from mc_dagprop_b200 import capi


def test_monte_carlo_class_name_is_primary():
    """reference test/test_naming_conventions.py:6-8"""
    assert MonteCarloPropagator.__name__ == "MonteCarloPropagator"
    assert Simulator is MonteCarloPropagator


def test_module_layout_matches_reference():
    from mc_dagprop.core import Activity as A2
    from mc_dagprop.monte_carlo import _core
    from mc_dagprop.types import ActivityIndex, ActivityType, EventId, EventIndex, ProbabilityMass, Second  # noqa: F401

    assert A2 is Activity and _core.MonteCarloPropagator is MonteCarloPropagator
    for name in ("GenericDelayGenerator", "DagContext", "SimResult", "Event", "Activity", "Simulator",
                 "MonteCarloPropagator", "EventTimestamp"):
        assert name in mc_dagprop.__all__ and hasattr(mc_dagprop, name)
    assert SimResult.__name__ == "SimResult"


def test_value_types_are_frozen_dataclasses():
    ts = EventTimestamp(earliest=1.0, latest=2.0, actual=3.0)
    ev = Event(event_id="a", timestamp=ts)
    act = Activity(idx=0, minimal_duration=1.0, activity_type=1)
    ctx = DagContext(events=[ev], activities={(0, 0): act}, precedence_list=[], max_delay=5.0)
    for obj in (ts, ev, act, ctx):
        assert dataclasses.is_dataclass(obj)
    with pytest.raises(dataclasses.FrozenInstanceError):
        act.idx = 3
    assert repr(act) == "Activity(idx=0, minimal_duration=1.0, activity_type=1)"
    assert repr(ts) == "EventTimestamp(earliest=1.0, latest=2.0, actual=3.0)"
    assert ev.timestamp.earliest == 1.0 and ctx.max_delay == 5.0 and ctx.activities[(0, 0)].idx == 0


def test_generator_api_and_errors():
    g = GenericDelayGenerator()
    g.set_seed(seed=3)
    g.add_constant(activity_type=1, factor=1.0)
    g.add_exponential(1, 1000.0, max_scale=1.0)
    g.add_exponential(activity_type=1, lambda_=2.0, max_scale=0.5)
    g.add_gamma(activity_type=1, shape=2.0, scale=1.0)
    g.add_gamma(activity_type=1, shape=2.0, scale=1.0, max_scale=0.5)
    g.add_empirical_absolute(activity_type=1, values=[10, 20], weights=[0.5, 0.5])
    g.add_empirical_relative(activity_type=1, factors=[1.0, 2.0], weights=[0.5, 0.5])
    with pytest.raises(RuntimeError, match="same length"):
        g.add_empirical_absolute(1, [1.0], [1.0, 2.0])
    with pytest.raises(RuntimeError, match="same length"):
        g.add_empirical_relative(1, [1.0], [1.0, 2.0])


def _ctx(max_delay=1e6, prec=None):
    events = [Event("0", EventTimestamp(0.0, 100.0, 0.0)), Event("1", EventTimestamp(0.0, 100.0, 0.0))]
    acts = {(0, 1): Activity(0, 1.0, 1), (1, 0): Activity(1, 1.0, 1)}
    return DagContext(events=events, activities=acts, precedence_list=prec or [(1, [(0, 0)])], max_delay=max_delay)


def test_constructor_errors_are_runtime_errors():
    """reference test/test_monte_carlo_extra.py:60-91 -- validation happens on the host, before any device work."""
    g = GenericDelayGenerator()
    g.add_constant(1, 0.0)
    with pytest.raises(RuntimeError, match="cycle"):
        Simulator(_ctx(prec=[(1, [(0, 0)]), (0, [(1, 1)])]), g)
    with pytest.raises(RuntimeError, match="max_delay"):
        Simulator(_ctx(max_delay=-1.0), g)
    g2 = GenericDelayGenerator()
    g2.add_constant(-1, 0.0)
    with pytest.raises(RuntimeError, match="reserved"):
        Simulator(_ctx(), g2)
    with pytest.raises(TypeError):
        Simulator(_ctx())  # generator is required


@pytest.mark.skipif(capi.device_count() > 0, reason="only meaningful without a GPU")
def test_no_silent_cpu_fallback():
    g = GenericDelayGenerator()
    g.add_constant(1, 0.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        Simulator(_ctx(), g)
