"""Executable-level golden: the results/continuous tree the UNMODIFIED reference (oracle/_ref/is3d_ref) writes for
BASELINE.json config 5 in miniature -- the first cells of the benchmark surface as a MUSIC-format (mode 6) surface.dat
(reference src/cpp/readindata.cpp:372-567), all 444 SMASH species, df_mode 2 with bulk + shear + baryon diffusion.

    python tests/golden/make_golden_exe_tree.py
tests/golden/exe_tree_music.npz holds the numeric content of the five result files of every 8th species (file order of
PDG/chosen_particles.dat) exactly as the reference wrote them (oracle/_ref prints 17 digits), the MC ids of ALL species and
the cell count; the test regenerates the same surface.dat from the seeds (synthetic.bench_surface + write_mode6)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import refrun  # noqa: E402
import subprocess  # noqa: E402
from is3d2_b200 import synthetic, workdir  # noqa: E402

FILES = ("dN_pTdpTdphidy", "vn", "dN_2pipTdpTdy", "dN_dphidy", "dN_dy")


def read_result(path):
    """numeric rows of a result file (header and blank lines dropped)"""
    rows = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            try:
                rows.append([float(v) for v in t])
            except ValueError:
                continue                      # header
    return np.array(rows)


def main():
    case = cases.EXE_TREE_CASE
    surf = cases.make_surface(case["surface"])
    with tempfile.TemporaryDirectory() as d:
        workdir.make_workdir(d, case["params"], chosen=case["chosen"])
        synthetic.write_mode6(os.path.join(d, "input", "surface.dat"), surf, baryon=True)
        with open(os.path.join(d, "ref_stdout.log"), "w") as log:
            r = subprocess.run([refrun.REF_BIN], cwd=d, stdout=log, stderr=subprocess.STDOUT)
        # include_baryon = 1: the reference segfaults at exit AFTER writing its results (destructor frees never-built splines,
        # SURVEY.md 8a) -- accept that exit status when the files are there
        mcids = np.loadtxt(os.path.join(d, "PDG", "chosen_particles.dat"), ndmin=1).astype(np.int64)
        out = {"mcid": mcids, "cells": np.int64(len(surf["tau"])), "exit_status": np.int64(r.returncode)}
        for s in range(0, len(mcids), 8):
            for stem in FILES:
                out[f"{stem}_{mcids[s]}"] = read_result(os.path.join(d, "results", "continuous", f"{stem}_{mcids[s]}.dat"))
        n_files = len(os.listdir(os.path.join(d, "results", "continuous")))
    out["n_files"] = np.int64(n_files)
    path = os.path.join(HERE, "exe_tree_music.npz")
    np.savez_compressed(path, **out)
    print("reference exit status", r.returncode, "files", n_files, "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
