"""Golden vectors for df_mode 5 (PTMA) under the CHAIN-FREE initial-guess policy, from the UNMODIFIED reference.

The reference starts a cell's Newton solve from the previous cell's solution when there is one and from (T, 1, 1)
otherwise (reference src/cpp/MomentumSpectra.cpp:1288-1313): on a ONE-cell surface there is never a previous solution, so
a one-cell run of the unmodified reference IS the chain-free policy.  This script runs oracle/_ref/is3d_ref once per cell
of each case's surface and adds the per-cell arrays up in cell order; Cooper-Frye spectra are additive over cells
(MomentumSpectra.cpp:1617-1640), so the sum is what the reference would produce for the whole surface if every cell
started from (T, 1, 1) -- the production policy of this repository (famod_chain = 0), the only one that shards.

    python tests/golden/make_golden_m5_chainfree.py [case ...]
Each tests/golden/m5free_<case>.npz holds the surface columns as the mode-1 reader reconstructs them and the summed array."""
import os
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import refrun  # noqa: E402
from is3d2_b200 import synthetic  # noqa: E402


def one_cell(args):
    surf, i, case, baryon = args
    cell = {k: v[i:i + 1] for k, v in surf.items()}
    with tempfile.TemporaryDirectory() as d:
        r = refrun.run_ref(d, cell, case["params"], chosen=case["chosen"], baryon=baryon, **case.get("tables", {}))
    return r["spectra"]


def main():
    only = set(sys.argv[1:])
    for name, case in cases.M5_CHAINFREE_CASES.items():
        if only and name not in only:
            continue
        surf = cases.make_surface(case["surface"])
        n = len(surf["tau"])
        baryon = bool(case["params"].get("include_baryon", 0))
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            per_cell = list(ex.map(one_cell, [(surf, i, case, baryon) for i in range(n)]))
        total = np.zeros_like(per_cell[0])
        for a in per_cell:                      # cell order, plain FP64 adds: the reference's own accumulation order
            total += a
        seen = synthetic.roundtrip_mode1(surf, baryon=baryon)
        out = os.path.join(HERE, f"m5free_{name}.npz")
        np.savez_compressed(out, spectra=total, **{f"col_{k}": v for k, v in seen.items()})
        print(name, n, "one-cell reference runs", total.shape, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
