"""Checks the rational approximation used by fast_atan (csrc/aniso.cuh): atan(x) = x + x z P(z)/Q(z), z = x^2, |x| <= 0.66,
with the three-range reduction (x > tan(3pi/8): pi/2 - atan(1/x); x > 0.66: pi/4 + atan((x-1)/(x+1))) against 50-digit
arithmetic.  P, Q are the classic degree-4 / degree-5 coefficients of the Cephes library's atan.c (public domain)."""
import mpmath as mp
import numpy as np

mp.mp.dps = 50
P = [-8.750608600031904122785e-1, -1.615753718733365076637e1, -7.500855792314704667340e1, -1.228866684490136173410e2,
     -6.485021904942025371773e1]
Q = [2.485846490142306297962e1, 1.650270098316988542046e2, 4.328810604912902668951e2, 4.853903996359136964868e2,
     1.945506571482613964425e2]


def fast_atan(s):
    if s > 2.414213562373095:
        y, num, den = np.pi / 2, -1.0, s
    elif s > 0.66:
        y, num, den = np.pi / 4, s - 1.0, s + 1.0
    else:
        y, num, den = 0.0, s, 1.0
    x = num / den
    z = x * x
    p = (((P[0] * z + P[1]) * z + P[2]) * z + P[3]) * z + P[4]
    q = ((((z + Q[0]) * z + Q[1]) * z + Q[2]) * z + Q[3]) * z + Q[4]
    r = x * z * p / q + x
    if s > 2.414213562373095:
        r += 6.123233995736765886130e-17
    elif s > 0.66:
        r += 0.5 * 6.123233995736765886130e-17
    return y + r


rng = np.random.default_rng(1)
worst = 0
for s in np.concatenate([rng.uniform(0, 0.66, 20000), rng.uniform(0.66, 2.5, 20000), np.exp(rng.uniform(0, 12, 20000)), [1e-8, 1e-3]]):
    e = abs(mp.mpf(fast_atan(float(s))) / mp.atan(mp.mpf(float(s))) - 1)
    worst = max(worst, e)
print("max relative error of fast_atan:", mp.nstr(worst, 4))
