"""Golden spin-polarization arrays: the unmodified reference (oracle/_ref) run on mode-5 surfaces (CPU-VH columns + six
thermal-vorticity columns) for tests/cases.py POLZN_CASES.  Stored: St, Sx, Sy, Sn, Snorm relabelled to the spectra
layout (Ns, NpT, Nphi, Ny), the vorticity columns and the surface as the reader reconstructs it."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import refrun  # noqa: E402
from is3d2_b200 import synthetic  # noqa: E402


def main():
    only = set(sys.argv[1:])
    for name, case in cases.POLZN_CASES.items():
        if only and name not in only:
            continue
        surf = cases.make_surface(case["surface"])
        with tempfile.TemporaryDirectory() as d:
            r = refrun.run_ref(d, surf, case["params"], chosen=case["chosen"], baryon=False, **case.get("tables", {}))
            flat = np.loadtxt(os.path.join(d, "input", "surface.dat"), ndmin=2)
            # the four result files as the reference writes them (4th column = S_mu / Snorm)
            files = np.stack([np.loadtxt(os.path.join(d, "results", f"S{c}.dat"), ndmin=2)[:, 3] for c in "txyn"])
        vort = np.ascontiguousarray(flat[:, 20:26].T)                    # as the reader parses them back
        seen = synthetic.roundtrip_mode1(surf, baryon=False)
        out = os.path.join(HERE, f"{name}.npz")
        if len(seen["tau"]) > 2000:      # large case: inputs are regenerated from their seeds in the test (synthetic.s3d + write_mode5)
            np.savez_compressed(out, polarization=r["polarization"])
        else:
            np.savez_compressed(out, polarization=r["polarization"], vorticity=vort, files=files, **{f"col_{k}": v for k, v in seen.items()})
        print(name, r["polarization"].shape, f"{r['seconds']:.1f}s", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
