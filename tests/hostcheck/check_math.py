"""TEST INFRASTRUCTURE (development aid): run the host-compiled kernel math (libhostcheck.so) against the reference
(oracle/_ref) on a few seeded surfaces.  Usage: python tests/hostcheck/check_math.py [df|feqmod]"""
import ctypes as C
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import refrun  # noqa: E402
from is3d_b200 import synthetic  # noqa: E402

hc = C.CDLL(os.path.join(HERE, "libhostcheck.so"))
hc.hostcheck_spectra_df.restype = C.c_long
hc.hostcheck_spectra_df.argtypes = [C.c_char_p, C.c_void_p, C.c_long]
hc.hostcheck_spectra_feqmod.restype = C.c_long
hc.hostcheck_spectra_feqmod.argtypes = [C.c_char_p, C.c_void_p, C.c_long, C.c_void_p]


def run(kind, name, s, P, chosen, baryon=False, **tb):
    root = "/tmp/hc_" + name
    r = refrun.run_ref(root, s, P, chosen=chosen, baryon=baryon, **tb)
    ref = r["spectra"]
    out = np.zeros(ref.size)
    st = np.zeros(4, dtype=np.int64)
    if kind == "df":
        hc.hostcheck_spectra_df(root.encode(), out.ctypes.data, out.size)
    else:
        hc.hostcheck_spectra_feqmod(root.encode(), out.ctypes.data, out.size, st.ctypes.data)
    out = out.reshape(ref.shape)
    m = np.abs(ref) > 1e-200 * np.abs(ref).max()
    rel = np.abs(out[m] / ref[m] - 1)
    log = open(root + "/ref_stdout.log").read()
    bd = re.findall(r"breaks down for (\d+)", log)
    pl = re.findall(r"pl went negative for (\d+)", log)
    print(f"{name:22s} {str(ref.shape):18s} max rel {rel.max():.3e} median {np.median(rel):.1e}  mine skip/break/pl {st[:3]}  ref break/pl {bd} {pl}")


B = dict(operation=1, mode=1, hrg_eos=2, dimension=3, include_baryon=0)
what = sys.argv[1] if len(sys.argv) > 1 else "feqmod"
if what == "feqmod":
    s = synthetic.s3d(200, seed=12345, stress=0.3)
    run("feqmod", "m3_3d", s, dict(B, df_mode=3), "pikp")
    run("feqmod", "m4_3d", s, dict(B, df_mode=4), "pikp")
    run("feqmod", "m3_3d_reg_out", s, dict(B, df_mode=3, hrg_eos=1, regulate_deltaf=1, outflow=1, deta_min=0.01), "pikp", phi_table="phi_table_48pt.dat")
    run("feqmod", "m4_3d_nobulk", s, dict(B, df_mode=4, include_bulk_deltaf=0), "pikp")
    run("feqmod", "m3_3d_noshear", s, dict(B, df_mode=3, include_shear_deltaf=0), "pikp")
    sb = synthetic.s3d(200, seed=7, baryon=True, stress=0.3)
    run("feqmod", "m3_3d_b", sb, dict(B, df_mode=3, include_baryon=1, include_baryondiff_deltaf=1), "pikp", baryon=True)
    run("feqmod", "m3_3d_b0", sb, dict(B, df_mode=3, include_baryon=1, include_baryondiff_deltaf=0), "pikp", baryon=True)
    s2 = synthetic.s3d(100, seed=3, dimension=2, stress=0.3)
    run("feqmod", "m3_2d", s2, dict(B, df_mode=3, dimension=2, hrg_eos=1), "pikp")
    run("feqmod", "m4_2d", s2, dict(B, df_mode=4, dimension=2, hrg_eos=3), "box", phi_table="phi_table_48pt.dat")
else:
    s = synthetic.s3d(200, seed=12345)
    run("df", "m1_3d", s, dict(B, df_mode=1), "pikp")
    run("df", "m2_3d", s, dict(B, df_mode=2), "pikp")
