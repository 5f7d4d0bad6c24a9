"""Times the pieces of the end-to-end call sequence (set_surface, spectra) separately: python tools/e2e_probe.py mode cells"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from is3d2_b200 import HostSession, synthetic, workdir
mode, cells = int(sys.argv[1]), int(sys.argv[2])
os.environ["IS3D_FAMOD_CHAIN"] = "0"
surf = synthetic.s3d(cells, seed=2024, baryon=True)
root = tempfile.mkdtemp()
workdir.make_workdir(root, bench.bench_params(mode), chosen="smash")
h = HostSession(root)
h.set_surface({k: v[:1000] for k, v in surf.items()})
h.prepare()
for rep in range(3):
    t0 = time.perf_counter(); h.abi_set_surface(surf); t1 = time.perf_counter()
    spec, st = h.abi_spectra(); t2 = time.perf_counter()
    print(f"mode {mode} cells {cells}: set_surface {1e3*(t1-t0):.1f} ms, spectra {1e3*(t2-t1):.1f} ms (kernel {st.kernel_ms:.1f} ms)")
