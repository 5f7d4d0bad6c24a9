"""Constants of fast_exp (is3d2_b200/csrc/common.cuh): N/ln2 and ln2/N (N = 1024), plus an error scan of the whole
construction (one-fma reduction with ln2/N rounded to double, degree-3 polynomial, table entry rounded to double) in
60-digit arithmetic, reported per range of x (the reduction error grows like |x| 1.1e-16)."""
import mpmath as mp
import numpy as np

N = 1024
mp.mp.dps = 60
L = mp.log(2) / N
step = float(L)
inv = float(N / mp.log(2))
print(f"{N}/ln2 =", repr(inv))
print(f"ln2/{N} =", repr(step))
tab = [float(mp.mpf(2) ** (mp.mpf(j) / N)) for j in range(N)]
rng = np.random.default_rng(0)
for lo, hi in ((-20, 10), (10, 30), (30, 60), (60, 200), (200, 707)):
    worst = 0
    for x in rng.uniform(lo, hi, 8000):
        k = int(np.rint(float(x) * inv))
        r = mp.mpf(float(mp.mpf(float(x)) - mp.mpf(k) * mp.mpf(step)))      # one fma: exact product, one rounding
        q = ((mp.mpf(1) / 6 * r + mp.mpf(1) / 2) * r + 1) * r
        T = mp.mpf(tab[k % N])
        v = (T * q + T) * mp.mpf(2) ** (k // N)
        worst = max(worst, abs(v / mp.exp(mp.mpf(float(x))) - 1))
    print(f"x in [{lo}, {hi}]: max relative error {mp.nstr(worst, 4)}")
