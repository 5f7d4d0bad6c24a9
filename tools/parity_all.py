"""Worst relative deviation from the unmodified reference's golden vectors, per spectra / dN/dX case (GPU box):
python tools/parity_all.py  -> one line per case + the overall worst (tests/harness.py tolerance rules)."""
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402

import cases  # noqa: E402
import harness  # noqa: E402

worst_all = 0.0
for name, case in cases.SPECTRA_CASES.items():
    surf, ref = harness.load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        with harness.open_session(tmp, case, surf) as h:
            got, st = h.abi_spectra()
    worst = harness.assert_spectra_close(got, ref, what=name)
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    print(f"{name:36s} worst {worst:.2e}  median {np.median(rel):.1e}  95th pct {np.percentile(rel, 95):.1e}")
    worst_all = max(worst_all, worst)
print(f"overall worst over {len(cases.SPECTRA_CASES)} spectra cases: {worst_all:.2e} (tolerance {harness.RTOL:g})")
