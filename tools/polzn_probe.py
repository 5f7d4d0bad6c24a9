"""Times the spin-polarization path (K7) on a synthetic mode-5 surface: python tools/polzn_probe.py cells [chosen]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from is3d2_b200 import HostSession, synthetic, workdir  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
chosen = sys.argv[2] if len(sys.argv) > 2 else "smash"
params = dict(operation=1, mode=5, hrg_eos=2, dimension=3, df_mode=2, include_baryon=0)
surf = synthetic.s3d(cells, seed=2024)
vort = np.random.default_rng(1).uniform(-0.05, 0.05, (6, cells))
root = tempfile.mkdtemp()
workdir.make_workdir(root, params, chosen=chosen)
h = HostSession(root)
h.set_surface({k: v[:1000] for k, v in surf.items()})
h.prepare()
h.abi_set_surface(surf)
h.abi_set_vorticity(vort)
shape = h.spectra_shape()
for rep in range(3):
    t0 = time.perf_counter()
    out, st = h.abi_polarization()
    dt = time.perf_counter() - t0
    evals = float(cells) * np.prod(shape)
    print(f"polarization, {cells} cells, {shape[0]} species: {dt * 1e3:.1f} ms (kernel {st.kernel_ms:.1f} ms) -> {evals / dt:.3e} evals/s")
