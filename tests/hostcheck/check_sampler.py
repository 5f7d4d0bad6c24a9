"""TEST INFRASTRUCTURE (development aid): host-compiled sampler math vs the reference's golden histograms."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
from is3d_b200 import synthetic, workdir  # noqa: E402

hc = C.CDLL(os.path.join(HERE, "libhostcheck.so"))
hc.hostcheck_sampler.restype = C.c_long
hc.hostcheck_sampler.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

GOLDEN = os.path.join(os.path.dirname(HERE), "golden")
for name, case in cases.SAMPLER_CASES.items():
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    surf = {k[4:]: z[k] for k in z.files if k.startswith("col_")}
    baryon = bool(case["params"].get("include_baryon", 0))
    root = workdir.make_workdir("/tmp/hcs_" + name, case["params"], chosen=case["chosen"])
    synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), surf, baryon=baryon)
    ns = z["dN_dy"].shape[0]
    nev = int(z["nevents"]) // 4
    dy, de, dp = np.zeros((ns, 100)), np.zeros((ns, 140)), np.zeros((ns, 100))
    yl = C.c_double(); prop = C.c_long()
    acc = hc.hostcheck_sampler(root.encode(), nev, dy.ctypes.data, de.ctypes.data, dp.ctypes.data, C.byref(yl), C.byref(prop))
    ref_scale = nev / int(z["nevents"])
    def chi2(a, b, sa, sb):   # a: mine (nev events), b: ref
        a = a.ravel(); b = b.ravel().astype(float)
        m = (a + b) > int(os.environ.get("THR", "20"))
        d = a[m] / sa - b[m] / sb
        v = a[m] / sa**2 + b[m] / sb**2
        return (d * d / v).sum(), int(m.sum())
    print(f"{name:20s} yield mine {yl.value:.12g} ref {float(z['total_yield']):.12g} rel {yl.value/float(z['total_yield'])-1:+.2e} | "
          f"accepted/event mine {acc/nev:.4f} ref {z['dN_deta'].sum()/int(z['nevents']):.4f} | prop {prop.value}")
    for key, mine in (("dN_dy", dy), ("dN_deta", de), ("dN_pT", dp)):
        c2, ndf = chi2(mine, z[key], nev, int(z["nevents"]))
        print(f"     {key:8s} chi2/ndf = {c2:9.1f} / {ndf}   per-species totals mine {mine.sum(axis=1)[:3]/nev} ref {z[key].sum(axis=1)[:3]/int(z['nevents'])}")
