#!/usr/bin/env python
"""BASELINE.json config 5 through the drop-in executable: THE benchmark surface (synthetic.bench_surface) written as a
MUSIC-format (mode 6) input/surface.dat, then iS3D_b200.e on IS3D_DEVICES (default: all GPUs of the box), phase times printed.

    python tools/exe_config5.py [cells=10000000] [devices=all]
The text file is ~670 bytes per cell (6.7 GB for 10 M cells); it is written block-parallel with pyarrow's CSV writer."""
import os
import shutil
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ProcessPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

from is3d2_b200 import synthetic, workdir  # noqa: E402

HBARC = synthetic.HBARC


def write_block(args):
    path, begin, end = args
    import pyarrow as pa
    import pyarrow.csv as pc
    s = synthetic.bench_surface(begin, end, baryon=True)
    tau = s["tau"]
    ut = np.sqrt(1.0 + s["ux"] ** 2 + s["uy"] ** 2 + (tau * s["un"]) ** 2)
    z = np.zeros_like(tau)
    E, T, P = s["E"] / HBARC, s["T"] / HBARC, s["P"] / HBARC
    cols = [tau, s["x"], s["y"], s["eta"], s["dat"] / tau, s["dax"] / tau, s["day"] / tau, s["dan"] / tau,
            ut, s["ux"], s["uy"], tau * s["un"], E, T, s["muB"] / HBARC, z, z, (E + P) / T,
            z, z, z, z, s["pixx"] / HBARC, s["pixy"] / HBARC, tau * s["pixn"] / HBARC, s["piyy"] / HBARC,
            tau * s["piyn"] / HBARC, z, s["bulkPi"] / HBARC, s["nB"], z, s["Vx"], s["Vy"], tau * s["Vn"]]   # synthetic.write_mode6
    tab = pa.table({f"c{i}": c for i, c in enumerate(cols)})
    pc.write_csv(tab, path, write_options=pc.WriteOptions(include_header=False, delimiter=" "))
    return path


def main():
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    devices = sys.argv[2] if len(sys.argv) > 2 else "all"
    params = dict(operation=1, mode=6, hrg_eos=2, dimension=3, df_mode=2, include_baryon=1, include_bulk_deltaf=1,
                  include_shear_deltaf=1, include_baryondiff_deltaf=1, regulate_deltaf=0, outflow=0)
    root = tempfile.mkdtemp(prefix="is3d_config5_")
    try:
        workdir.make_workdir(root, params, chosen="smash")
        t0 = time.time()
        step = 625_000
        jobs = [(os.path.join(root, f"part_{k:04d}.dat"), b, min(b + step, cells)) for k, b in enumerate(range(0, cells, step))]
        with ProcessPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            parts = list(ex.map(write_block, jobs))
        target = os.path.join(root, "input", "surface.dat")
        with open(target, "wb") as out:
            for p in parts:
                with open(p, "rb") as f:
                    shutil.copyfileobj(f, out, 64 << 20)
                os.remove(p)
        print(f"surface.dat: {cells} cells, MUSIC layout, {os.path.getsize(target) / 1e9:.2f} GB written in {time.time() - t0:.1f} s")
        exe = os.path.join(REPO, "is3d2_b200", "iS3D_b200.e")
        env = dict(os.environ, IS3D_TIMING="1", IS3D_READER_VERBOSE="1", IS3D_DEVICES=devices)
        env.pop("IS3D_DEVICE", None)
        t0 = time.time()
        r = subprocess.run([exe], cwd=root, capture_output=True, text=True, env=env)
        wall = time.time() - t0
        keep = [l for l in r.stdout.splitlines() if l.startswith(("[timing]", "[reader]", "Sharding", "Number of freezeout", "Spectra calculation", "Finished"))]
        print("\n".join(keep))
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
        n_files = len(os.listdir(os.path.join(root, "results", "continuous")))
        print(f"exit {r.returncode}; executable wall time {wall:.2f} s for {cells} cells on IS3D_DEVICES={devices}; {n_files} result files")
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
