"""Fits the polynomial of the FP64 atan2 of the kernels (hmp_kernels.cu, atan2_r(double, double)):
atan(t) = t * Q(t^2) on |t| <= tan(pi/8), Q of degree N by Chebyshev interpolation in 60-digit arithmetic, then checks the
double-precision Horner evaluation against mpmath on a dense grid. Prints the coefficients as C literals.
usage: python tools/fit_atan_f64.py [degree]"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60
N = int(sys.argv[1]) if len(sys.argv) > 1 else 13
umax = mp.tan(mp.pi / 8) ** 2 * mp.mpf("1.0001")


def Q(u):
    if u == 0:
        return mp.mpf(1)
    s = mp.sqrt(u)
    return mp.atan(s) / s


# Chebyshev nodes on [0, umax], interpolation polynomial in the monomial basis of u (solved in high precision)
n = N + 1
nodes = [(umax / 2) * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * n))) for k in range(n)]
A = mp.matrix(n, n)
b = mp.matrix(n, 1)
for i, u in enumerate(nodes):
    for j in range(n):
        A[i, j] = u ** j
    b[i] = Q(u)
c = mp.lu_solve(A, b)
coef = [float(c[j]) for j in range(n)]
coef[0] = 1.0


def atan_poly(t):
    u = t * t
    p = coef[N]
    for j in range(N - 1, 0, -1):
        p = p * u + coef[j]
    return t + t * (u * p)   # coef[0] = 1 exactly


worst = 0.0
for t in np.linspace(0.0, float(mp.tan(mp.pi / 8)), 20001):
    ref = mp.atan(mp.mpf(float(t)))
    err = abs(mp.mpf(atan_poly(float(t))) - ref)
    rel = float(err / ref) if ref != 0 else 0.0
    worst = max(worst, rel)
print(f"// degree {N} in t^2, max relative error of the double evaluation on [0, tan(pi/8)]: {worst:.3e}")
for j in range(1, n):
    print(f"\t{coef[j]!r},")
