"""Structural checks of an ``AnalyticContext`` before it is handed to the GPU engine.

The reference validates in ``analytic/_context.py:93-158`` and reports every violation as ``ValueError``; the same
conditions are checked here (grid step, ``max_delay``, event windows, activities, precedence entries, acyclicity),
grouped into one function per kind of object.
"""
from __future__ import annotations

import numpy as np


def _check_scalars(context) -> None:
    if context.step <= 0.0:
        raise ValueError("step_size must be positive")
    cap = context.max_delay
    if cap is not None and cap < 0.0:
        raise ValueError("max_delay must be non-negative when provided")


def _check_events(context) -> None:
    for number, event in enumerate(context.events):
        window = event.timestamp
        if window.latest < window.earliest:
            raise ValueError(f"event {number} has earliest > latest")
        if window.actual < window.earliest or window.actual > window.latest:
            raise ValueError(f"event {number} actual time outside bounds")


def _check_activities(context) -> None:
    count = len(context.events)
    for edge_key, (_, activity) in context.activities.items():
        if any(node < 0 or node >= count for node in edge_key):
            raise ValueError(f"activity {edge_key} references invalid node")
        pmf = activity.pmf
        pmf.validate()
        if not np.isclose(pmf.step, context.step):
            raise ValueError(f"edge {edge_key} step {pmf.step} does not match context step size {context.step}")
        pmf.validate_alignment(context.step)
        mass = pmf.total_mass
        if not np.isclose(mass, 1.0):
            raise ValueError(f"activity {edge_key} PMF does not sum to 1, got {mass}")


def _check_precedence(context) -> None:
    """Index ranges, edge bookkeeping, and a cycle check by peeling events whose predecessors are all placed."""
    count = len(context.events)
    waiting = np.zeros(count, dtype=np.int64)
    fan_out: dict[int, list[int]] = {}
    for target, predecessors in context.precedence_list:
        if target < 0 or target >= count:
            raise ValueError(f"target index {target} out of range")
        for source, activity_index in predecessors:
            if source < 0 or source >= count:
                raise ValueError(f"predecessor index {source} out of range")
            known = context.activities.get((source, target))
            if known is None:
                raise ValueError(f"missing activity for {(source, target)}")
            if known[0] != activity_index:
                raise ValueError(
                    f"edge index {activity_index} for {(source, target)} does not match context mapping {known[0]}")
            fan_out.setdefault(source, []).append(target)
            waiting[target] += 1
    frontier = [int(i) for i in np.flatnonzero(waiting == 0)]
    placed = 0
    while frontier:
        placed += 1
        for follower in fan_out.get(frontier.pop(), ()):
            waiting[follower] -= 1
            if waiting[follower] == 0:
                frontier.append(follower)
    if placed != count:
        raise ValueError("precedence list contains a cycle")


def validate_context(context) -> None:
    """Raise ``ValueError`` for the first structural problem of ``context``."""
    _check_scalars(context)
    _check_events(context)
    _check_activities(context)
    _check_precedence(context)
