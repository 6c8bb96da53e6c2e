"""Type stubs of ``mc_dagprop_b200.monte_carlo._core`` (also reachable as ``mc_dagprop.monte_carlo``).

First half: the surface of the reference module (``src/mc_dagprop/monte_carlo/_core.pyi:1-101``,
bound in ``_core.cpp:366-552``) -- same class names, field names and keyword names, so code typed
against WonJayne/mc_dagprop keeps checking.  Second half: the additive API of the B200 engine
(array ingest, array-shaped results, fused statistics, duration injection, device selection).
"""
from collections.abc import Collection, Iterable, Mapping, Sequence
from typing import TypedDict

import numpy as np
from numpy.typing import ArrayLike, NDArray

from mc_dagprop_b200.types import ActivityIndex, ActivityType, EventId, EventIndex, Second

# ---------------------------------------------------------------------------------------------
# reference surface
# ---------------------------------------------------------------------------------------------

class EventTimestamp:
    """Time window of an event.  The Monte-Carlo engine reads ``earliest`` only (``_core.cpp:319,333``)."""

    earliest: Second
    latest: Second
    actual: Second
    def __init__(self, earliest: Second, latest: Second, actual: Second) -> None: ...

class Event:
    """A node of the DAG: identifier plus time window (frozen dataclass, ``_core.cpp:446-488``)."""

    event_id: EventId
    timestamp: EventTimestamp
    def __init__(self, event_id: EventId, timestamp: EventTimestamp) -> None: ...

class Activity:
    """An edge of the DAG.  ``idx`` is the row of ``SimResult.durations`` the activity reports to;
    ``activity_type`` selects the delay distribution (types without one keep ``minimal_duration``)."""

    idx: ActivityIndex
    minimal_duration: Second
    activity_type: ActivityType
    def __init__(self, idx: ActivityIndex, minimal_duration: Second, activity_type: ActivityType) -> None: ...

class DagContext:
    """Events, activities, precedence list (any order; sorted topologically by the propagator, a
    cycle raises ``RuntimeError``) and the cap ``realized <= earliest + max_delay``."""

    events: Sequence[Event]
    activities: Mapping[tuple[EventIndex, EventIndex], Activity]
    precedence_list: Sequence[tuple[EventIndex, list[tuple[EventIndex, ActivityIndex]]]]
    max_delay: Second
    def __init__(
        self,
        events: Sequence[Event],
        activities: Mapping[tuple[EventIndex, EventIndex], Activity],
        precedence_list: Sequence[tuple[EventIndex, Sequence[tuple[EventIndex, ActivityIndex]]]],
        max_delay: Second,
    ) -> None: ...

class SimResult:
    """One sample.  The arrays are writeable zero-copy views whose ``.base`` is this object
    (``_core.cpp:491-516``); ``numpy.asarray(result)`` is ``realized`` (buffer protocol).  All results
    of one ``run_many`` call view one shared batch, which lives until the last of them is released."""

    @property
    def realized(self) -> NDArray[np.float64]: ...
    @property
    def durations(self) -> NDArray[np.float64]: ...
    @property
    def cause_event(self) -> NDArray[np.int32]: ...
    def __buffer__(self, flags: int, /) -> memoryview: ...

class GenericDelayGenerator:
    """Per-``activity_type`` delay distributions (``_core.cpp:146-159``); a later ``add_*`` for the
    same type replaces the earlier one; type ``-1`` is reserved."""

    def __init__(self) -> None: ...
    def set_seed(self, seed: int) -> None: ...
    def add_constant(self, activity_type: ActivityType, factor: float) -> None: ...
    def add_exponential(self, activity_type: ActivityType, lambda_: float, max_scale: float) -> None: ...
    def add_gamma(
        self, activity_type: ActivityType, shape: float, scale: float, max_scale: float = ...
    ) -> None: ...
    def add_empirical_absolute(
        self, activity_type: ActivityType, values: Collection[Second], weights: Collection[float]
    ) -> None: ...
    def add_empirical_relative(
        self, activity_type: ActivityType, factors: Collection[Second], weights: Collection[float]
    ) -> None: ...

# ---------------------------------------------------------------------------------------------
# additive API of the B200 engine
# ---------------------------------------------------------------------------------------------

class ReducedStats(TypedDict, total=False):
    """Result of :meth:`MonteCarloPropagator.run_many_reduced` (statistics of ``realized - earliest``)."""

    n: int
    sum: NDArray[np.float64]              # [E]
    sumsq: NDArray[np.float64]            # [E]
    late: NDArray[np.uint64]              # [len(thresholds), E]  counts of delay > threshold
    hist: NDArray[np.uint32]              # [E, n_bins]
    cause_activity: NDArray[np.uint64]    # [A]  only with cause_counts=True
    cause_none: NDArray[np.uint64]        # [E]  only with cause_counts=True

class MonteCarloPropagator:
    """Monte-Carlo propagator on B200.  Reference surface: the two-argument constructor, ``run``,
    ``run_many``, ``node_count``, ``activity_count`` (``_core.cpp:545-551``).  ``device`` /
    ``devices`` choose the CUDA device(s); with several devices every call shards its seeds into
    contiguous blocks, one per device, and returns exactly what one device would return."""

    def __init__(
        self,
        context: DagContext,
        generator: GenericDelayGenerator,
        device: int = 0,
        devices: Sequence[int] | None = None,
    ) -> None: ...
    def node_count(self) -> int: ...
    def activity_count(self) -> int: ...
    def run(self, seed: int) -> SimResult: ...
    def run_many(self, seeds: Iterable[int]) -> list[SimResult]: ...

    # -- additive --
    @staticmethod
    def from_arrays(
        earliest: ArrayLike,
        act_idx: ArrayLike,
        act_base: ArrayLike,
        act_type: ArrayLike,
        prec_target: ArrayLike,
        prec_off: ArrayLike,
        pred_src: ArrayLike,
        pred_act: ArrayLike,
        max_delay: float,
        generator: GenericDelayGenerator,
        device: int = 0,
        devices: Sequence[int] | None = None,
    ) -> MonteCarloPropagator: ...
    def level_count(self) -> int: ...
    def device(self) -> int: ...
    def devices(self) -> list[int]: ...
    def set_option(self, option: int, value: int) -> None: ...
    def run_many_arrays(
        self, seeds: ArrayLike
    ) -> tuple[NDArray[np.float64], NDArray[np.float64], NDArray[np.int32]]: ...
    def run_with_durations(self, durations: ArrayLike) -> tuple[NDArray[np.float64], NDArray[np.int32]]: ...
    def run_many_reduced(
        self,
        seeds: ArrayLike,
        thresholds: Sequence[float] = ...,
        n_bins: int = 0,
        hist_lo: float = 0.0,
        hist_hi: float = 1.0,
        cause_counts: bool = False,
    ) -> ReducedStats: ...

def set_pinned_cache_limit(bytes: int) -> None: ...
def pinned_cache_bytes() -> int: ...
