"""Times the continuous-spectra path on a synthetic surface in any geometry:
    python tools/spectra_probe.py df_mode dimension cells [phi_table] [chosen]
e.g. `2 2 200000 phi_table_48pt.dat` = boost-invariant surface, 24 eta nodes x 48 phi x 51 pT, all SMASH species."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from is3d2_b200 import HostSession, synthetic, workdir  # noqa: E402

mode, dim, cells = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
phi = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None
chosen = sys.argv[5] if len(sys.argv) > 5 else "smash"
params = dict(bench.bench_params(mode), dimension=dim)
baryon = bool(params["include_baryon"])
surf = synthetic.s3d(cells, seed=2024, baryon=baryon, dimension=dim)
root = tempfile.mkdtemp()
workdir.make_workdir(root, params, chosen=chosen, phi_table=phi)
os.environ.setdefault("IS3D_FAMOD_CHAIN", "0")
h = HostSession(root)
h.set_surface({k: v[:1000] for k, v in surf.items()})
h.prepare()
ns, npT, nphi, ny = h.spectra_shape()
neta = 1 if dim == 3 else len(np.loadtxt(os.path.join(root, "tables", "spacetime_rapidity", "eta_table.dat")))
h.abi_set_surface(surf)
for rep in range(3):
    t0 = time.perf_counter()
    out, st = h.abi_spectra()
    dt = time.perf_counter() - t0
    evals = float(cells) * ns * npT * nphi * ny * neta
    print(f"spectra df_mode {mode}, {dim}+1d, {cells} cells, {ns} species, {npT} pT x {nphi} phi x {ny} y x {neta} eta: "
          f"{dt * 1e3:.1f} ms (kernels {st.kernel_ms:.1f} ms) -> {evals / dt:.3e} evals/s")
