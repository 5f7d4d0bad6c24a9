"""Generate the committed golden vectors by running the UNMODIFIED reference (oracle/_ref/is3d_ref, built by
oracle/Makefile from /root/reference) on the seeded cases of tests/cases.py.  Run in the build container:
    python tests/golden/make_golden.py
Each tests/golden/<case>.npz holds the surface columns exactly as the mode-1 reader reconstructs them and the
reference's in-memory dN_pTdpTdphidy array (binary dump, full double precision)."""
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
    # the launch-realistic case (BIG_SPECTRA_CASES, ~90 s of serial reference) only when named on the command line
    todo = dict(cases.SPECTRA_CASES)
    todo.update({n: c for n, c in cases.BIG_SPECTRA_CASES.items() if n in only})
    for name, case in todo.items():
        if only and name not in only:
            continue
        out = os.path.join(HERE, f"spectra_{name}.npz")
        surf = cases.make_surface(case["surface"])
        baryon = bool(case["params"].get("include_baryon", 0))
        with tempfile.TemporaryDirectory() as d:
            r = refrun.run_ref(d, surf, case["params"], chosen=case["chosen"], baryon=baryon, **case.get("tables", {}))
        seen = synthetic.roundtrip_mode1(surf, baryon=baryon)
        np.savez_compressed(out, spectra=r["spectra"], **{f"col_{k}": v for k, v in seen.items()})
        print(name, r["spectra"].shape, f"{r['seconds']:.2f}s", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
