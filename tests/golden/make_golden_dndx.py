"""Golden vectors of the spacetime distributions dN/dX: the reference's results/continuous files (written with 17
digits by oracle/_ref, see oracle/ref_precision.h) for the seeded cases of tests/cases.py DNDX_CASES.
Stored exactly as the reference wrote them: normalised, and with its species-cumulative partial-memset behaviour
(SpacetimeDistribution.cpp:166-168)."""
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
    todo = dict(cases.DNDX_CASES)
    todo.update({n: c for n, c in cases.BIG_DNDX_CASES.items() if n in only})      # the large case only when named
    for name, case in todo.items():
        if only and name not in only:
            continue
        surf = cases.make_surface(case["surface"])
        baryon = bool(case["params"].get("include_baryon", 0))
        with tempfile.TemporaryDirectory() as d:
            refrun.run_ref(d, surf, case["params"], chosen=case["chosen"], baryon=baryon, **case.get("tables", {}))
            mcids = np.loadtxt(os.path.join(d, "PDG", "chosen_particles.dat"), ndmin=1)
            h = refrun.read_dndx_files(d, mcids)
        seen = synthetic.roundtrip_mode1(surf, baryon=baryon)
        out = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(out, tau=h["tau"], r=h["r"], phi=h["phi"], **{f"col_{k}": v for k, v in seen.items()})
        print(name, h["tau"].shape, h["r"].shape, h["phi"].shape, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
