"""Discretised delay distributions for the analytic propagator (the builders of reference
``analytic/distributions.py``: ``constant_pmf``, ``exponential_pmf``, ``gamma_pmf``, ``empirical_pmf``).

A continuous law is put on the grid by differencing its CDF at the grid edges ``start, start + step, ...`` and
normalising over the kept range: bin ``k`` carries the mass of ``[edge_k, edge_k+1)``.  Host-side numpy; the PMFs
these functions return are what the GPU engine consumes.
"""
from __future__ import annotations

import math
from typing import Callable, Iterable

import numpy as np

from ..types import Second
from ._pmf import DiscretePMF


def _on_grid(pmf: DiscretePMF, step) -> DiscretePMF:
    pmf.validate()
    pmf.validate_alignment(step)
    return pmf


def _discretise(cdf: Callable[[np.ndarray], np.ndarray], step, start, stop) -> DiscretePMF:
    """Masses of the bins between consecutive grid edges in ``[start, stop]``, normalised to one."""
    if step <= 0.0:
        raise ValueError("step must be positive")
    if stop < start:
        raise ValueError("stop must be greater or equal to start")
    edges = np.arange(start, stop + step, step)
    mass = np.diff(cdf(edges))
    total = mass.sum()
    if total == 0.0:
        raise ValueError("zero probability mass in range")
    return _on_grid(DiscretePMF(edges[:-1], mass / total, step=step), step)


def _gamma_p(shape: float, x: np.ndarray, rel_tol: float = 1e-12, max_terms: int = 200) -> np.ndarray:
    """Regularised lower incomplete gamma function P(shape, x), elementwise: the power series
    ``x^a e^-x / Gamma(a) * sum_n x^n / (a (a+1) ... (a+n))`` where ``x < shape + 1``, one minus the modified-Lentz
    continued fraction of Q(shape, x) elsewhere (Abramowitz & Stegun 6.5.29 / 6.5.31)."""
    x = np.asarray(x, dtype=float)
    if shape <= 0.0 or np.any(x < 0.0):
        raise ValueError("shape must be > 0 and x >= 0")
    out = np.zeros_like(x)
    positive = x > 0.0
    with np.errstate(divide="ignore"):
        front = np.where(positive, np.exp(-x + shape * np.log(np.where(positive, x, 1.0)) - math.lgamma(shape)), 0.0)
    low = positive & (x < shape + 1.0)
    if np.any(low):
        xs = x[low]
        term = np.full_like(xs, 1.0 / shape)
        acc = term.copy()
        denom = shape
        for _ in range(max_terms):
            denom += 1.0
            term = term * xs / denom
            acc += term
            if np.all(np.abs(term) < np.abs(acc) * rel_tol):
                break
        out[low] = acc * front[low]
    high = positive & ~low
    if np.any(high):
        xs = x[high]
        floor = 1e-30
        b = xs + 1.0 - shape
        c = np.full_like(xs, 1.0 / floor)
        d = 1.0 / b
        frac = d.copy()
        for n in range(1, max_terms + 1):
            a_n = -n * (n - shape)
            b = b + 2.0
            d = a_n * d + b
            d = np.where(np.abs(d) < floor, floor, d)
            c = b + a_n / c
            c = np.where(np.abs(c) < floor, floor, c)
            d = 1.0 / d
            step_factor = d * c
            frac = frac * step_factor
            if np.all(np.abs(step_factor - 1.0) < rel_tol):
                break
        out[high] = 1.0 - front[high] * frac
    return out


def constant_pmf(value: Second, step: int) -> DiscretePMF:
    """All mass at ``value``."""
    return _on_grid(DiscretePMF.delta(value, step), step)


def exponential_pmf(scale: Second, step: int, start: int, stop: int) -> DiscretePMF:
    """Exponential with mean ``scale`` on ``[start, stop)``."""
    if scale <= 0.0:
        raise ValueError("scale must be positive")
    return _discretise(lambda t: 1.0 - np.exp(-t / scale), step, start, stop)


def gamma_pmf(shape: float, scale: Second, step: int, start: Second, stop: Second) -> DiscretePMF:
    """Gamma(shape, scale) on ``[start, stop)``."""
    if shape <= 0.0 or scale <= 0.0:
        raise ValueError("shape and scale must be positive")
    return _discretise(lambda t: _gamma_p(shape, t / scale), step, start, stop)


def empirical_pmf(values: Iterable[Second], weights: Iterable[float], step: int) -> DiscretePMF:
    """``values`` with relative ``weights``."""
    support = np.array(list(values), dtype=float)
    weight = np.array(list(weights), dtype=float)
    if support.size != weight.size:
        raise ValueError("values and weights must have same length")
    total = weight.sum()
    if total <= 0.0:
        raise ValueError("weights must sum to a positive number")
    return _on_grid(DiscretePMF(support, weight / total, step=step), step)
