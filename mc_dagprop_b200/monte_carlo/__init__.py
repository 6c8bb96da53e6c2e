"""Drop-in Monte-Carlo classes (reference ``mc_dagprop/monte_carlo/__init__.py:3-5``)."""
from __future__ import annotations

try:
    from ._core import (
        Activity,
        DagContext,
        Event,
        EventTimestamp,
        GenericDelayGenerator,
        MonteCarloPropagator,
        SimResult,
    )
except ModuleNotFoundError as exc:  # pragma: no cover - compiled module missing
    raise ImportError(
        "mc_dagprop_b200 requires the compiled extension 'mc_dagprop_b200.monte_carlo._core' and "
        "libmcdp_b200.so; build them with `python -m mc_dagprop_b200.build` (there is no CPU fallback)."
    ) from exc

Simulator = MonteCarloPropagator

__all__ = [
    "GenericDelayGenerator",
    "DagContext",
    "SimResult",
    "Event",
    "Activity",
    "MonteCarloPropagator",
    "Simulator",
    "EventTimestamp",
]
