"""Worst relative deviation from the unmodified reference's golden vectors, per continuous-spectra case (GPU box):
python tools/parity_all.py  -> one line per case + the overall worst (tests/harness.py tolerance rules).  Covers the
multi-cell goldens (df_mode 5 under the reference's serial chain), the launch-realistic 2304-cell SMASH case and the
df_mode 5 chain-free goldens (sums of one-cell reference runs; the library's default policy)."""
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402

import cases  # noqa: E402
import harness  # noqa: E402


def run(name, case, surf, ref, famod_chain):
    with tempfile.TemporaryDirectory() as tmp:
        with harness.open_session(tmp, case, surf, famod_chain=famod_chain) as h:
            got, st = h.abi_spectra()
    worst = harness.assert_spectra_close(got, ref, what=name)
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    print(f"{name:44s} worst {worst:.2e}  median {np.median(rel):.1e}  95th pct {np.percentile(rel, 95):.1e}")
    return worst


worst_all, n = 0.0, 0
for name, case in list(cases.SPECTRA_CASES.items()) + list(cases.BIG_SPECTRA_CASES.items()):
    surf, ref = harness.load_golden(name)
    worst_all = max(worst_all, run(name, case, surf, ref, 1))
    n += 1
for name, case in cases.M5_CHAINFREE_CASES.items():
    surf, ref = harness.load_golden_m5free(name)
    worst_all = max(worst_all, run("m5 chain-free (one-cell reference runs): " + name, case, surf, ref, 0))
    n += 1
print(f"overall worst over {n} spectra cases: {worst_all:.2e} (tolerance {harness.RTOL:g})")
