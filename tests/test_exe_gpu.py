"""Executable-level end to end (BASELINE.json config 5 in miniature): a MUSIC-format (mode 6) surface.dat -> iS3D_b200.e ->
results/continuous tree, compared file by file with the tree the UNMODIFIED reference wrote for the same input
(tests/golden/make_golden_exe_tree.py; reference readers readindata.cpp:372-567, writers EmissionFunction.cpp:406-558, :804-878)."""
import os
import subprocess

import numpy as np
import pytest

import cases
from is3d2_b200 import synthetic, workdir

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exe_tree_music.npz")
FILES = ("dN_pTdpTdphidy", "vn", "dN_2pipTdpTdy", "dN_dphidy", "dN_dy")


def _rows(path):
    rows = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            try:
                rows.append([float(v) for v in t])
            except ValueError:
                continue
    return np.array(rows)


@pytest.mark.parametrize("devices", [None, "all"])
def test_executable_results_tree_matches_the_reference_tree(libs, tmp_path, devices):
    z = np.load(GOLDEN)
    case = cases.EXE_TREE_CASE
    surf = cases.make_surface(case["surface"])
    assert len(surf["tau"]) == int(z["cells"])
    root = workdir.make_workdir(str(tmp_path), case["params"], chosen=case["chosen"])
    synthetic.write_mode6(os.path.join(root, "input", "surface.dat"), surf, baryon=True)
    exe = os.path.join(workdir.REPO, "is3d2_b200", "iS3D_b200.e")
    env = dict(os.environ)
    env.pop("IS3D_DEVICE", None)
    if devices:
        env["IS3D_DEVICES"] = devices             # every GPU of the box: cells sharded inside the executable
    r = subprocess.run([exe], cwd=root, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "Finished particlization" in r.stdout
    out = os.path.join(root, "results", "continuous")
    mcid = z["mcid"]
    names = sorted(os.listdir(out))
    assert len(names) == int(z["n_files"]) == 5 * len(mcid)
    assert names == sorted(f"{stem}_{m}.dat" for stem in FILES for m in mcid)
    worst = 0.0
    for s in range(0, len(mcid), 8):
        for stem in FILES:
            ref = z[f"{stem}_{mcid[s]}"]
            got = _rows(os.path.join(out, f"{stem}_{mcid[s]}.dat"))
            assert got.shape == ref.shape, (stem, mcid[s], got.shape, ref.shape)
            # our files carry the reference's own 9 significant digits (setprecision(8) scientific; oracle/_ref prints 17):
            # half a unit of the last printed digit is 5e-9 relative; bins that are cancellations get 1e-12 of the column peak
            peak = np.abs(ref).max(axis=0, keepdims=True)
            err = np.abs(got - ref)
            tol = 6e-9 * np.abs(ref) + 1e-12 * peak + 1e-300
            if stem == "dN_dy":
                tol = 6e-8 * np.abs(ref) + 1e-300     # this writer prints 8 significant digits (setprecision(8), not scientific)
            if stem == "vn":
                tol = tol + 2e-8          # flow coefficients are ratios of sums, printed with 9 digits
            assert np.all(err <= tol), (stem, int(mcid[s]), float((err / tol).max()))
            worst = max(worst, float((err / (np.abs(ref) + 1e-12 * peak + 1e-300)).max()))
    print(f"results tree vs reference ({devices or 'one GPU'}): worst relative deviation {worst:.2e} (print precision 5e-9)")
