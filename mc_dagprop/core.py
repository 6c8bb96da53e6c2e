from mc_dagprop_b200.core import Activity, DagContext, Event, EventTimestamp  # noqa: F401

__all__ = ["EventTimestamp", "Event", "Activity", "DagContext"]
