"""Discretised delay distributions for the analytic propagator (reference ``analytic/distributions.py``):
the mass of ``[edge_k, edge_k+1)`` from the CDF at the grid edges, normalised over the kept range."""
from __future__ import annotations

import math
from typing import Iterable

import numpy as np

from ..types import Second
from ._pmf import DiscretePMF


def _checked(pmf: DiscretePMF, step) -> DiscretePMF:
    pmf.validate()
    pmf.validate_alignment(step)
    return pmf


def constant_pmf(value: Second, step: int) -> DiscretePMF:
    """All mass at ``value``."""
    return _checked(DiscretePMF.delta(value, step), step)


def _lower_regularized_gamma(shape: float, x: float) -> float:
    """P(shape, x): power series below ``shape + 1``, Lentz continued fraction of Q above (Numerical Recipes 6.2)."""
    if shape <= 0.0 or x < 0.0:
        raise ValueError("shape must be > 0 and x >= 0")
    if x == 0.0:
        return 0.0
    log_front = -x + shape * math.log(x) - math.lgamma(shape)
    if x < shape + 1.0:
        term = total = 1.0 / shape
        a = shape
        for _ in range(200):
            a += 1.0
            term *= x / a
            total += term
            if abs(term) < abs(total) * 1e-12:
                break
        return total * math.exp(log_front)
    tiny = 1e-30
    b = x + 1.0 - shape
    c, d = 1.0 / tiny, 1.0 / b
    h = d
    for i in range(1, 201):
        an = -i * (i - shape)
        b += 2.0
        d = an * d + b
        d = tiny if abs(d) < tiny else d
        c = b + an / c
        c = tiny if abs(c) < tiny else c
        d = 1.0 / d
        delta = d * c
        h *= delta
        if abs(delta - 1.0) < 1e-12:
            break
    return 1.0 - math.exp(log_front) * h


def _masses_from_cdf(cdf: np.ndarray) -> np.ndarray:
    mass = np.diff(cdf)
    total = mass.sum()
    if total == 0.0:
        raise ValueError("zero probability mass in range")
    return mass / total


def _range_checks(step, start, stop) -> None:
    if step <= 0.0:
        raise ValueError("step must be positive")
    if stop < start:
        raise ValueError("stop must be greater or equal to start")


def exponential_pmf(scale: Second, step: int, start: int, stop: int) -> DiscretePMF:
    """Exponential with mean ``scale``, discretised on ``[start, stop)``."""
    if scale <= 0.0:
        raise ValueError("scale must be positive")
    _range_checks(step, start, stop)
    edges = np.arange(start, stop + step, step)
    return _checked(DiscretePMF(edges[:-1], _masses_from_cdf(1.0 - np.exp(-edges / scale)), step=step), step)


def gamma_pmf(shape: float, scale: Second, step: int, start: Second, stop: Second) -> DiscretePMF:
    """Gamma(shape, scale), discretised on ``[start, stop)``."""
    if shape <= 0.0 or scale <= 0.0:
        raise ValueError("shape and scale must be positive")
    _range_checks(step, start, stop)
    edges = np.arange(start, stop + step, step)
    cdf = np.array([_lower_regularized_gamma(shape, e / scale) for e in edges])
    return _checked(DiscretePMF(edges[:-1], _masses_from_cdf(cdf), step=step), step)


def empirical_pmf(values: Iterable[Second], weights: Iterable[float], step: int) -> DiscretePMF:
    """``values`` with relative ``weights``."""
    v = np.array(list(values), dtype=float)
    w = np.array(list(weights), dtype=float)
    if v.size != w.size:
        raise ValueError("values and weights must have same length")
    if w.sum() <= 0.0:
        raise ValueError("weights must sum to a positive number")
    return _checked(DiscretePMF(v, w / w.sum(), step=step), step)
