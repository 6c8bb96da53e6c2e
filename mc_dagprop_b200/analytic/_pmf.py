"""``DiscretePMF``: a probability mass function on an equidistant grid (reference ``analytic/_pmf.py``).

Same fields, checks and methods as the reference class; ``convolve`` and ``maximum`` run on the GPU
(``mcdp_pmf_op``, compensated twice-working-precision accumulation in place of the reference's ``np.longdouble``) and apply the reference's
mass correction (``_expected_mass`` / ``_rescale``, ``_pmf.py:80-105``) there.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..types import ProbabilityMass, Second
from . import _device


@dataclass(frozen=True, slots=True)
class DiscretePMF:
    values: np.ndarray
    probabilities: np.ndarray
    step: int  # grid spacing of ``values`` (integer seconds)

    def __post_init__(self) -> None:
        self.validate()
        if self.step < 0.0:
            raise ValueError("step size must be non-negative")
        if not isinstance(self.step, int):
            raise OverflowError(
                f"step must be an integer number of seconds, got: {self.step} (type: {type(self.step)})"
                "we limit to ints to avoid floating point precision issues")

    # -- checks (reference _pmf.py:38-66) ------------------------------------------------------------
    def validate(self) -> None:
        n = len(self.values)
        if n == 0:
            raise ValueError("PMF values cannot be empty")
        if n != len(self.probabilities):
            raise ValueError("values and probs must have same length")
        if n > 1 and np.any(np.diff(self.values) < 0):
            raise ValueError("values must be sorted in non-decreasing order")
        mass = self.probabilities.sum()
        if mass > 1.0 and not np.isclose(mass, 1.0):
            raise ValueError("Probabilities must sum to <= 1.0")

    def validate_alignment(self, step: Second) -> None:
        if not np.isclose(self.step, step):
            raise ValueError(f"PMF step {self.step} does not match expected {step}")
        if step <= 0.0:
            raise ValueError("step must be positive")
        if len(self.values) == 0:
            raise ValueError("PMF values cannot be empty")
        if len(self.values) > 1 and not np.allclose(np.diff(self.values), step):
            raise ValueError("PMF grid spacing does not match step")
        if not np.isclose(self.values[0] % step, 0.0):
            raise ValueError("PMF values are not aligned to step grid")

    # -- construction helpers ------------------------------------------------------------------------
    @staticmethod
    def _unchecked(values: np.ndarray, probabilities: np.ndarray, step: int) -> "DiscretePMF":
        """For results of the device engine whose checks were done in bulk (``AnalyticPropagator.run``)."""
        pmf = object.__new__(DiscretePMF)
        object.__setattr__(pmf, "values", values)
        object.__setattr__(pmf, "probabilities", probabilities)
        object.__setattr__(pmf, "step", step)
        return pmf

    @staticmethod
    def delta(v: Second, step: Second) -> "DiscretePMF":
        """A unit mass at ``v``."""
        return DiscretePMF(np.array([v], dtype=float), np.array([1.0], dtype=float), step=step)

    @property
    def total_mass(self) -> ProbabilityMass:
        return ProbabilityMass(self.probabilities.sum())

    def shift(self, delta: Second) -> "DiscretePMF":
        return DiscretePMF(self.values + delta, self.probabilities.copy(), step=self.step)

    # -- device operations ----------------------------------------------------------------------------
    def _grid(self) -> tuple[int, np.ndarray]:
        return int(round(float(self.values[0]))), np.ascontiguousarray(self.probabilities, np.float64)

    def _from_device(self, start: int, probs: np.ndarray) -> "DiscretePMF":
        values = float(start) + float(self.step) * np.arange(len(probs), dtype=float)
        return DiscretePMF(values, probs, step=self.step)

    def convolve(self, other: "DiscretePMF") -> "DiscretePMF":
        """PMF of the sum of two independent variables (reference ``_pmf.py:107-125``), on the GPU."""
        a0, pa = self._grid()
        b0, pb = other._grid()
        start, probs, _, _ = _device.pmf_op(0, self.step, a0, pa, b0, pb)
        if len(pa) == 1 and len(pb) > 1:  # a delta shifts the other operand: its own value grid is kept
            return DiscretePMF(other.values + self.values[0], probs, step=self.step)
        if len(pb) == 1:
            return DiscretePMF(self.values + other.values[0], probs, step=self.step)
        return self._from_device(start, probs)

    def maximum(self, other: "DiscretePMF") -> "DiscretePMF":
        """PMF of ``max(X, Y)`` for independent variables (reference ``_pmf.py:127-148``), on the GPU."""
        a0, pa = self._grid()
        b0, pb = other._grid()
        start, probs, _, _ = _device.pmf_op(1, self.step, a0, pa, b0, pb)
        return self._from_device(start, probs)
