"""Times surface.dat ingestion (host layer) for 1 thread and all hardware threads:
python tools/reader_probe.py [cells]"""
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from is3d2_b200 import HostSession, synthetic, workdir  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
s = synthetic.s3d(n, seed=5, baryon=True)
root = tempfile.mkdtemp()
workdir.make_workdir(root, dict(hrg_eos=2, df_mode=2, dimension=3, mode=1, include_baryon=1), chosen="pikp")
p = os.path.join(root, "input", "surface.dat")
np.savetxt(p, np.stack([s[k] for k in synthetic.SOA_COLUMNS], axis=1), fmt="%.17g")
print(f"{n} cells, {os.path.getsize(p) / 1e6:.0f} MB of text, {os.cpu_count()} hardware threads", file=sys.stderr)
os.environ["IS3D_READER_VERBOSE"] = "1"
for nt in (1, os.cpu_count() or 1):
    os.environ["IS3D_READER_THREADS"] = str(nt)
    with HostSession(root) as h:
        t0 = time.time()
        h.read_surface()
        dt = time.time() - t0
    print(f"{nt} threads: {dt:.2f} s -> {n / dt / 1e6:.2f} M cells/s, {os.path.getsize(p) / dt / 1e6:.0f} MB/s", file=sys.stderr)
shutil.rmtree(root)
