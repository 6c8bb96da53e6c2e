"""Compatibility layer re-exporting the core dataclasses (reference ``mc_dagprop/core.py``)."""
from __future__ import annotations

from .monte_carlo import Activity, DagContext, Event, EventTimestamp

__all__ = ["EventTimestamp", "Event", "Activity", "DagContext"]
