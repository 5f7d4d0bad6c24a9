"""Constants of fast_exp (is3d_b200/csrc/common.cuh): N/ln2 and the two-part split of ln2/N (N = 1024) whose high part
has 26 significant bits (so k * hi is exact for |k| < 2^27), plus an error scan of the whole construction (reduction,
degree-3 polynomial, table entry rounded to double) in 60-digit arithmetic."""
import mpmath as mp
import numpy as np

N = 1024
mp.mp.dps = 60
L = mp.log(2) / N
m, e = np.frexp(float(L))
hi = float(np.ldexp(np.floor(m * 2 ** 26) / 2 ** 26, e))      # keep 26 significant bits
lo = float(L - mp.mpf(hi))
inv = float(N / mp.log(2))
print(f"{N}/ln2 =", repr(inv))
print("hi =", repr(hi), " lo =", repr(lo))
tab = [float(mp.mpf(2) ** (mp.mpf(j) / N)) for j in range(N)]
rng = np.random.default_rng(0)
worst = 0
for x in np.concatenate([rng.uniform(-20, 60, 20000), rng.uniform(60, 707, 2000)]):
    k = int(np.rint(float(x) * inv))
    r = mp.mpf(float(x)) - mp.mpf(k) * mp.mpf(hi) - mp.mpf(k) * mp.mpf(lo)
    r = mp.mpf(float(r))                                       # r is rounded to double once (second fma)
    q = ((mp.mpf(1) / 6 * r + mp.mpf(1) / 2) * r + 1) * r
    T = mp.mpf(tab[k % N])
    v = (T * q + T) * mp.mpf(2) ** (k // N)
    worst = max(worst, abs(v / mp.exp(mp.mpf(float(x))) - 1))
print("max relative error (exact arithmetic on the rounded constants):", mp.nstr(worst, 4))
