"""End-to-end run of the drop-in executable iS3D_b200.e on a synthetic surface.dat: python tools/exe_probe.py cells [operation] [df_mode]"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from is3d2_b200 import synthetic, workdir  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
operation = int(sys.argv[2]) if len(sys.argv) > 2 else 1
df_mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
baryon = 0 if df_mode == 4 else 1
params = dict(operation=operation, mode=1, hrg_eos=2, dimension=3, df_mode=df_mode, include_baryon=baryon,
              include_baryondiff_deltaf=baryon, oversample=1, fast=1, test_sampler=0, sampler_seed=1,
              min_num_hadrons=1.0e7, max_num_samples=1000)
s = synthetic.s3d(n, seed=5, baryon=bool(baryon), stress=0.3 if df_mode >= 3 else 0.0)
root = tempfile.mkdtemp()
workdir.make_workdir(root, params, chosen="smash")
cols = synthetic.SOA_COLUMNS if baryon else synthetic.SOA_COLUMNS[:20]
a = np.stack([s[k] for k in cols], axis=1)
a[:, 11:20] /= 0.197327053                    # file columns are in fm^-1 units
if baryon:
    a[:, 20] /= 0.197327053
np.savetxt(os.path.join(root, "input", "surface.dat"), a, fmt="%.17g")
exe = os.path.join(workdir.REPO, "is3d2_b200", "iS3D_b200.e")
env = dict(os.environ, IS3D_READER_VERBOSE="1", IS3D_FAMOD_CHAIN="0")
t0 = time.time()
r = subprocess.run([exe], cwd=root, capture_output=True, text=True, env=env)
dt = time.time() - t0
lines = [l for l in r.stdout.splitlines() if l.strip()]
print("\n".join(lines[-25:]))
print(f"exit {r.returncode}; {n} cells, operation {operation}, df_mode {df_mode}: wall {dt:.2f} s; results: "
      f"{len(os.listdir(os.path.join(root, 'results', 'continuous')))} continuous files, "
      f"{len([f for f in os.listdir(os.path.join(root, 'results')) if f.startswith('particle_list')])} particle lists")
shutil.rmtree(root)
