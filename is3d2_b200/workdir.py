"""Build a runnable iS3D working directory (the reference and the drop-in executable both use fixed relative
paths: iS3D_parameters.dat, input/surface.dat, PDG/, tables/, deltaf_coefficients/, results/ --
reference src/cpp/iS3D.cpp:97,233,254-257; clear_results.sh:3-14)."""
from __future__ import annotations

import os
import re
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(REPO, "data")

RESULT_DIRS = ["results/continuous", "results/sampled/vn", "results/sampled/dN_taudtaudy",
               "results/sampled/dN_2pirdrdy", "results/sampled/dN_dphisdy", "results/sampled/dN_2pipTdpTdy",
               "results/sampled/dN_dphipdy", "results/sampled/dN_dy", "results/sampled/dN_deta"]


def default_parameters() -> dict:
    """key -> value string of the shipped iS3D_parameters.dat (order preserved)."""
    out = {}
    with open(os.path.join(DATA, "iS3D_parameters.dat")) as f:
        for line in f:
            line = line.split("#")[0]
            if "=" in line:
                k, v = line.split("=", 1)
                out[k.strip()] = v.strip()
    return out


def write_parameters(path: str, overrides: dict) -> dict:
    p = default_parameters()
    for k, v in overrides.items():
        if k not in p:
            raise KeyError(f"unknown iS3D parameter {k!r}")
        p[k] = repr(v) if isinstance(v, float) else str(v)
    with open(path, "w") as f:
        for k, v in p.items():
            f.write(f"{k} = {v}\n")
    return p


def make_workdir(root: str, params: dict, chosen: str = "pikp", phi_table: str | None = None,
                 pT_table: str | None = None, y_table: str | None = None) -> str:
    """Create `root` with data copies, parameter file and results tree.  `chosen` names a
    PDG/chosen_particles_<chosen>.dat list (or is a path)."""
    os.makedirs(root, exist_ok=True)
    for d in ("PDG", "tables"):
        dst = os.path.join(root, d)
        if os.path.exists(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(DATA, d), dst)
    os.makedirs(os.path.join(root, "tables", "thermodynamic"), exist_ok=True)
    link = os.path.join(root, "deltaf_coefficients")
    if not os.path.exists(link):
        os.symlink(os.path.join(DATA, "deltaf_coefficients"), link)
    src = chosen if os.path.exists(chosen) else os.path.join(DATA, "PDG", f"chosen_particles_{chosen}.dat")
    shutil.copyfile(src, os.path.join(root, "PDG", "chosen_particles.dat"))
    mom = os.path.join(root, "tables", "momentum")
    if phi_table:
        shutil.copyfile(os.path.join(DATA, "tables", "momentum", phi_table), os.path.join(mom, "phi_table.dat"))
    if pT_table:
        shutil.copyfile(os.path.join(DATA, "tables", "momentum", pT_table), os.path.join(mom, "pT_table.dat"))
    if y_table:
        shutil.copyfile(os.path.join(DATA, "tables", "momentum", y_table), os.path.join(mom, "y_table.dat"))
    os.makedirs(os.path.join(root, "input"), exist_ok=True)
    for d in RESULT_DIRS:
        os.makedirs(os.path.join(root, d), exist_ok=True)
    write_parameters(os.path.join(root, "iS3D_parameters.dat"), params)
    return root
