"""Exact-arithmetic restatement of the operation order behind ``np.linalg.inv`` for 4x4 matrices (OpenBLAS dgesv).

TEST INFRASTRUCTURE.  ``inverse4_emulated`` performs every floating-point operation with a correctly rounded FMA built
from ``fractions.Fraction`` (slow, exact), in the order the device function ``inverse4`` (msp_device.cuh) uses:
left-looking LU with dot products accumulated from zero, pivot = first largest magnitude, sub-diagonal scaled by the
reciprocal of the pivot, forward substitution ascending, back substitution DESCENDING with a reciprocal multiply.
tests/test_oracle_golden.py checks that this equals np.linalg.inv bit for bit, which pins the device order to NumPy.
"""
from fractions import Fraction as Fr

import numpy as np


def _fma(a, b, c):
    return float(Fr(a) * Fr(b) + Fr(c))


def inverse4_emulated(M):
    n = 4
    a = [[float(v) for v in row] for row in np.asarray(M, dtype=np.float64)]
    perm, ipiv = list(range(n)), list(range(n))
    for j in range(n):
        b = [a[i][j] for i in range(n)]
        for i in range(j):
            ip = ipiv[i]
            if ip != i:
                b[i], b[ip] = b[ip], b[i]
        for i in range(1, j):
            acc = 0.0
            for k in range(i):
                acc = _fma(a[i][k], b[k], acc)
            b[i] = b[i] - acc
        for r in range(j, n):
            acc = 0.0
            for k in range(j):
                acc = _fma(a[r][k], b[k], acc)
            b[r] = b[r] - acc
        jp = max(range(j, n), key=lambda r: (abs(b[r]), -r))
        ipiv[j] = jp
        if b[jp] == 0.0:
            raise np.linalg.LinAlgError("Singular matrix")
        if jp != j:
            b[j], b[jp] = b[jp], b[j]
            for k in range(j):
                a[j][k], a[jp][k] = a[jp][k], a[j][k]
            perm[j], perm[jp] = perm[jp], perm[j]
        rp = 1.0 / b[j]
        for r in range(j + 1, n):
            b[r] = b[r] * rp
        for i in range(n):
            a[i][j] = b[i]
    inv = np.zeros((n, n))
    for j in range(n):
        y = [0.0] * n
        for i in range(n):
            s = 1.0 if perm[i] == j else 0.0
            for k in range(i):
                s = _fma(-a[i][k], y[k], s)
            y[i] = s
        x = [0.0] * n
        for i in range(n - 1, -1, -1):
            s = y[i]
            for k in range(n - 1, i, -1):
                s = _fma(-a[i][k], x[k], s)
            x[i] = s * (1.0 / a[i][i])
        for i in range(n):
            inv[i][j] = x[i]
    return inv
