"""Golden histograms of the Monte-Carlo sampler: the reference (oracle/_ref) run with test_sampler = 1 on the seeded
cases of tests/cases.py SAMPLER_CASES.  Stored as integer counts per species and bin, the number of sampled events
and the reference's exact mean total yield (calculate_total_yield, dumped in binary by oracle/ref_harness.cpp)."""
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
    for name, case in cases.SAMPLER_CASES.items():
        if only and name not in only:
            continue
        surf = cases.make_surface(case["surface"])
        baryon = bool(case["params"].get("include_baryon", 0))
        with tempfile.TemporaryDirectory() as d:
            r = refrun.run_ref(d, surf, case["params"], chosen=case["chosen"], baryon=baryon, **case.get("tables", {}))
            mcids = np.loadtxt(os.path.join(d, "PDG", "chosen_particles.dat"), ndmin=1)
            h = refrun.read_sampler_test_files(d, mcids, case["params"])
        seen = synthetic.roundtrip_mode1(surf, baryon=baryon)
        out = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(out, **{k: (v.astype(np.int32) if isinstance(v, np.ndarray) else v) for k, v in h.items()},
                            **{f"col_{k}": v for k, v in seen.items()})
        print(name, "events", h["nevents"], "hadrons", int(h["dN_deta"].sum()), "mean yield/event", h["total_yield"],
              f"{r['seconds']:.1f}s", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
