"""Fits the polynomial of the FP64 exp of the kernels (hmp_kernels.cu, exp_fast_d): exp(r) = 1 + r (1 + r q(r)) on
|r| <= ln(2) / 2, q of degree N by Chebyshev interpolation in 60-digit arithmetic; checks the double-precision evaluation
against mpmath. Prints the coefficients of q (ascending) as C literals.   usage: python tools/fit_exp_f64.py [degree]"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60
N = int(sys.argv[1]) if len(sys.argv) > 1 else 9
rmax = mp.log(2) / 2 * mp.mpf("1.001")


def q(r):
    if r == 0:
        return mp.mpf(1) / 2
    return (mp.exp(r) - 1 - r) / (r * r)


n = N + 1
nodes = [rmax * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
A = mp.matrix(n, n)
b = mp.matrix(n, 1)
for i, r in enumerate(nodes):
    for j in range(n):
        A[i, j] = r ** j
    b[i] = q(r)
c = mp.lu_solve(A, b)
coef = [float(c[j]) for j in range(n)]


def exp_poly(r):
    p = coef[N]
    for j in range(N - 1, -1, -1):
        p = p * r + coef[j]
    return (p * r + 1.0) * r + 1.0


worst = 0.0
for r in np.linspace(-float(mp.log(2) / 2), float(mp.log(2) / 2), 20001):
    ref = mp.exp(mp.mpf(float(r)))
    worst = max(worst, float(abs(mp.mpf(exp_poly(float(r))) - ref) / ref))
print(f"// q of degree {N}, max relative error of the double evaluation of exp(r) on |r| <= ln2/2: {worst:.3e}")
for j in range(n):
    print(f"\t{coef[j]!r},")
