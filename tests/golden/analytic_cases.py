"""Seeded inputs of the analytic-propagator parity cases, shared by the fixture generator (which builds them with the
REFERENCE's classes) and the GPU tests (which build them with this repository's drop-in classes)."""
import numpy as np

RULES = (1, 2, 3)  # truncate, remove, redistribute

# (name, n_events, seed, step, max_delay or None, underflow rule, overflow rule)
CASES = [
    ("chain_tt", 12, 1, 1, None, 1, 1),
    ("dag_tt", 40, 2, 1, None, 1, 1),
    ("dag_rr", 40, 3, 1, 25, 2, 2),
    ("dag_dd", 40, 4, 1, 30, 3, 3),
    ("dag_tr", 60, 5, 1, 40, 1, 2),
    ("dag_dt", 60, 6, 1, None, 3, 1),
    ("dag_rd_step5", 30, 7, 5, 60, 2, 3),
    ("wide_tt", 25, 8, 1, None, 1, 1),
    ("tight_rr", 50, 9, 1, 6, 2, 2),
    ("tight_dd", 50, 10, 1, 6, 3, 3),
    ("flat_rr_b", 40, 11, 1, 40, 2, 2),
    ("flat_rt", 40, 12, 1, None, 2, 1),
    ("flat_tr", 40, 13, 1, 12, 1, 2),
    ("flat_rd", 40, 14, 1, 12, 2, 3),
]


def describe(n_events, seed, step, max_delay, wide=False, flat=False):
    """A layered random DAG: events with [earliest, latest] windows on the grid, 1-3 predecessors each, one PMF per
    edge with a support of a few to a few dozen bins (negative delays included).  `flat`: every predecessor is an
    origin event (the first third of the events) -- the reference's own mass assertion (`_propagator.py:139-145`)
    rejects a REMOVE rule that actually removed mass upstream of another bounded event."""
    rng = np.random.default_rng(1000 + seed)
    earliest = np.sort(rng.integers(0, 4 * n_events, size=n_events)) * step
    earliest[0] = 0
    latest = earliest + rng.integers(8, 60 if not wide else 400, size=n_events) * step
    events = [(f"e{i}", float(earliest[i]), float(latest[i]), float(earliest[i])) for i in range(n_events)]
    activities, precedence = {}, []
    idx = 0
    n_origin = max(1, n_events // 3)
    for tgt in range(n_origin if flat else 1, n_events):
        if rng.random() < 0.1:
            continue
        k = int(rng.integers(1, 4))
        srcs = sorted(set(rng.integers(0 if flat else max(0, tgt - 6), n_origin if flat else tgt, size=k).tolist()))
        preds = []
        for s in srcs:
            n_bins = int(rng.integers(1, 12 if not wide else 300))
            first = int(rng.integers(-3, 6)) * step + int(round((earliest[tgt] - earliest[s]) / step)) * step
            probs = rng.random(n_bins) ** 2 + 1e-3
            probs /= probs.sum()
            values = first + step * np.arange(n_bins, dtype=float)
            activities[(s, tgt)] = (idx, values, probs)
            preds.append((s, idx))
            idx += 1
        precedence.append((tgt, tuple(preds)))
    return events, activities, tuple(precedence)


def build_context(ns, name, n_events, seed, step, max_delay, under, over):
    """`ns`: anything with Event, EventTimestamp, DiscretePMF, AnalyticActivity, AnalyticContext, UnderflowRule,
    OverflowRule attributes (the reference's package or the drop-in)."""
    events, activities, precedence = describe(n_events, seed, step, max_delay, wide=name.startswith("wide"),
                                              flat=2 in (under, over))
    evs = tuple(ns.Event(i, ns.EventTimestamp(e, l, a)) for i, e, l, a in events)
    acts = {k: (idx, ns.AnalyticActivity(idx, ns.DiscretePMF(v, p, step=step))) for k, (idx, v, p) in activities.items()}
    return ns.AnalyticContext(events=evs, activities=acts, precedence_list=precedence, step=step,
                              underflow_rule=ns.UnderflowRule(under), overflow_rule=ns.OverflowRule(over), max_delay=max_delay)


def pmf_pairs(seed=0, n=24):
    """Operand pairs for DiscretePMF.convolve / maximum: deltas, short and long supports, partial masses, tiny tails."""
    rng = np.random.default_rng(77 + seed)
    out = []
    for i in range(n):
        la, lb = int(rng.integers(1, 6 if i % 3 else 200)), int(rng.integers(1, 6 if i % 4 else 150))
        a0, b0 = int(rng.integers(-20, 20)), int(rng.integers(-20, 20))
        pa, pb = rng.random(la) + 1e-6, rng.random(lb) ** 8 + 1e-300
        pa /= pa.sum()
        pb /= pb.sum()
        if i % 5 == 0:
            pa *= 0.7  # a truncated operand: masses multiply
        out.append((a0, pa, b0, pb))
    return out
