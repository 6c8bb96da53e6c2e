"""Coefficients of kLogG in csrc/mcdp_math.cuh: g(t) = sum_{k>=1} t^(k-1)/(2k+1), so that
log(m) = 2 s + 2 s^3 g(s^2) with s = (m-1)/(m+1); Chebyshev interpolation of degree 6 on
[0, ((sqrt2-1)/(sqrt2+1))^2] in 80-bit arithmetic."""
import numpy as np

ld = np.longdouble


def g(t):
    t = np.asarray(t, dtype=ld)
    acc = np.zeros_like(t)
    for k in range(60, 0, -1):
        acc = acc * t + ld(1) / ld(2 * k + 1)
    return acc


smax = (np.sqrt(ld(2)) - 1) / (np.sqrt(ld(2)) + 1)
tmax = float(smax * smax) * 1.0005
n = 7
k = np.arange(n, dtype=ld)
t = (np.cos(np.pi * (2 * k + 1) / (2 * n)) + 1) * ld(tmax) / 2
A = np.concatenate([np.vander(t, n, increasing=True).astype(ld), g(t)[:, None]], axis=1)
for i in range(n):
    p = np.argmax(np.abs(A[i:, i])) + i
    A[[i, p]] = A[[p, i]]
    A[i] = A[i] / A[i, i]
    for j in range(n):
        if j != i:
            A[j] = A[j] - A[j, i] * A[i]
print([float(v).hex() for v in A[:, -1]])
