"""Times the dN/dX path (K4) on a synthetic surface: python tools/dndx_probe.py df_mode cells [chosen]"""
import ctypes as C
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from is3d2_b200 import HostSession, Stats, synthetic, workdir  # noqa: E402

mode, cells = int(sys.argv[1]), int(sys.argv[2])
chosen = sys.argv[3] if len(sys.argv) > 3 else "smash"
params = dict(bench.bench_params(mode), operation=0)
baryon = bool(params["include_baryon"])
surf = synthetic.s3d(cells, seed=2024, baryon=baryon, stress=0.3 if mode >= 3 else 0.0)
root = tempfile.mkdtemp()
workdir.make_workdir(root, params, chosen=chosen)
h = HostSession(root)
h.set_surface({k: v[:1000] for k, v in surf.items()})
h.prepare()
ns = h.spectra_shape()[0]
npT, nphi, ny = h.spectra_shape()[1:]
tau, r, phi = np.zeros((ns, 120)), np.zeros((ns, 60)), np.zeros((ns, 100))
h.abi_set_surface(surf)
for rep in range(3):
    st = Stats()
    t0 = time.perf_counter()
    rc = h.lib.is3d_dndx(h.ctx, tau.ctypes.data, r.ctypes.data, phi.ctypes.data, C.byref(st))
    dt = time.perf_counter() - t0
    assert rc == 0, h.lib.is3d_last_error(h.ctx)
    evals = float(cells) * ns * npT * nphi * ny
    print(f"dN/dX df_mode {mode}, {cells} cells, {ns} species: {dt * 1e3:.1f} ms (kernels {st.kernel_ms:.1f} ms) -> {evals / dt:.3e} evals/s"
          f" (reruns {st.prune_reruns}, thread-slot evaluations dropped {st.evals_dropped:.3g})")
