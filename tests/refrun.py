"""TEST INFRASTRUCTURE: run the compiled reference (oracle/_ref/is3d_ref, built by oracle/Makefile from the
unmodified sources) in a scratch working directory and read back its binary dumps."""
from __future__ import annotations

import os
import struct
import subprocess

import numpy as np

from is3d2_b200 import synthetic, workdir

REPO = workdir.REPO
REF_BIN = os.path.join(REPO, "oracle", "_ref", "is3d_ref")
REF_BIN_OMP = os.path.join(REPO, "oracle", "_ref", "is3d_ref_omp")


def have_ref() -> bool:
    return os.access(REF_BIN, os.X_OK)


def run_ref(root: str, surface: dict, params: dict, chosen: str = "pikp", baryon: bool | None = None,
            omp_threads: int = 0, timeout: float = 3600.0, **tables) -> dict:
    if baryon is None:
        baryon = bool(int(params.get("include_baryon", 0)))
    workdir.make_workdir(root, params, chosen=chosen, **tables)
    if int(params.get("mode", 1)) == 5:       # mode 1 columns + thermal vorticity (seeded by the cell count)
        synthetic.write_mode5(os.path.join(root, "input", "surface.dat"), surface, baryon=baryon, seed=len(surface["tau"]))
    else:
        synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), surface, baryon=baryon)
    env = dict(os.environ)
    exe = REF_BIN
    if omp_threads:
        exe = REF_BIN_OMP
        env["OMP_NUM_THREADS"] = str(omp_threads)
    with open(os.path.join(root, "ref_stdout.log"), "w") as log:
        r = subprocess.run([exe], cwd=root, stdout=log, stderr=subprocess.STDOUT, env=env, timeout=timeout)
    if r.returncode != 0:
        tail = open(os.path.join(root, "ref_stdout.log")).read()[-2000:]
        raise RuntimeError(f"reference exited {r.returncode}:\n{tail}")
    return read_dumps(root)


def ref_surface(root: str, timeout: float = 600.0) -> np.ndarray:
    """Run only the reference's surface reader in `root` (parameters and input/surface.dat already in place) and return the
    parsed cells as an (n, 31) array in FO_surf field order (25 hot-path columns + 6 vorticity components)."""
    env = dict(os.environ, IS3D_REF_SURFACE_ONLY="1")
    with open(os.path.join(root, "ref_stdout.log"), "w") as log:
        r = subprocess.run([REF_BIN], cwd=root, stdout=log, stderr=subprocess.STDOUT, env=env, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference reader exited {r.returncode}:\n" + open(os.path.join(root, "ref_stdout.log")).read()[-2000:])
    raw = open(os.path.join(root, "ref_dump", "surface.bin"), "rb").read()
    n = struct.unpack("l", raw[:8])[0]
    return np.frombuffer(raw[8:], dtype=np.float64).reshape(n, 31).copy()


def read_dumps(root: str) -> dict:
    out = {}
    d = os.path.join(root, "ref_dump")
    p = os.path.join(d, "spectra.bin")
    if os.path.exists(p):
        raw = open(p, "rb").read()
        dims = struct.unpack("4l", raw[:32])
        out["spectra"] = np.frombuffer(raw[32:], dtype=np.float64).reshape(dims).copy()
    p = os.path.join(d, "polarization.bin")
    if os.path.exists(p):
        raw = open(p, "rb").read()
        ns, npT, nphi, ny = struct.unpack("4l", raw[:32])
        a = np.frombuffer(raw[32:], dtype=np.float64).reshape(5, ny, nphi, npT, ns)      # storage order: species fastest
        out["polarization"] = np.ascontiguousarray(np.transpose(a, (0, 4, 3, 2, 1)))       # -> (5, Ns, NpT, Nphi, Ny)
    p = os.path.join(d, "species.bin")
    if os.path.exists(p):
        raw = open(p, "rb").read()
        n = struct.unpack("l", raw[:8])[0]
        out["species"] = np.frombuffer(raw[8:], dtype=np.float64).reshape(n, 8).copy()
    p = os.path.join(d, "jonah.bin")
    if os.path.exists(p):
        raw = open(p, "rb").read()
        m = struct.unpack("l", raw[:8])[0]
        a = np.frombuffer(raw[8:], dtype=np.float64)
        out["jonah"] = {"bulkPi_over_P": a[:m].copy(), "lambda2": a[m:2 * m].copy(), "z": a[2 * m:3 * m].copy(),
                        "max": float(a[3 * m])}
    p = os.path.join(d, "particles.bin")
    if os.path.exists(p):
        raw = open(p, "rb").read()
        nev = struct.unpack("l", raw[:8])[0]
        off = 8
        events = []
        for _ in range(nev):
            npart = struct.unpack("l", raw[off:off + 8])[0]
            off += 8
            events.append(np.frombuffer(raw[off:off + npart * 104], dtype=np.float64).reshape(npart, 13).copy())
            off += npart * 104
        out["events"] = events
    p = os.path.join(d, "timing.txt")
    if os.path.exists(p):
        out["seconds"] = float(open(p).read().split()[0])
    return out


def read_dndx_files(root: str, mcids) -> dict:
    """results/continuous/{dN_taudtaudy,dN_2pirdrdy,dN_dphidy}_<mcid>.dat of a reference run (17 digits in
    oracle/_ref): arrays [species][bins] of the NORMALISED values exactly as written."""
    out = {}
    for key, stem in (("tau", "dN_taudtaudy"), ("r", "dN_2pirdrdy"), ("phi", "dN_dphidy")):
        rows = []
        for m in mcids:
            a = np.loadtxt(os.path.join(root, "results", "continuous", f"{stem}_{int(m)}.dat"), ndmin=2)
            rows.append(a[:, 1])
        out[key] = np.array(rows)
    return out


def read_sampler_test_files(root: str, mcids, params: dict) -> dict:
    """results/sampled/*_test.dat of a reference run with test_sampler = 1, converted back to integer COUNTS
    (writers: EmissionFunction.cpp:685-975).  Also Nevents and the exact mean total yield dumped by the harness."""
    from is3d2_b200 import workdir
    p = workdir.default_parameters()
    p.update({k: str(v) for k, v in params.items()})
    f = lambda k: float(p[k])  # noqa: E731
    nev = int(open(os.path.join(root, "ref_dump", "nevents.txt")).read())
    y_cut = f("y_cut")
    yw = 2.0 * y_cut / f("y_bins")
    ew = 2.0 * f("eta_cut") / f("eta_bins")
    pw = (f("pT_max") - f("pT_min")) / f("pT_bins")
    phw = 2.0 * np.pi / f("phip_bins")
    tw = (f("tau_max") - f("tau_min")) / f("tau_bins")
    rw = (f("r_max") - f("r_min")) / f("r_bins")
    out = {k: [] for k in ("dN_dy", "dN_deta", "dN_pT", "dN_dphip", "dN_tau", "dN_r", "dN_phis")}
    for m in mcids:
        m = int(m)
        def col(sub, stem):
            return np.loadtxt(os.path.join(root, "results", "sampled", sub, f"{stem}_{m}_test.dat"), ndmin=2)
        a = col("dN_dy", "dN_dy"); out["dN_dy"].append(a[:, 1] * yw * nev)
        a = col("dN_deta", "dN_deta"); out["dN_deta"].append(a[:, 1] * ew * nev)
        a = col("dN_2pipTdpTdy", "dN_2pipTdpTdy"); out["dN_pT"].append(a[:, 1] * (2.0 * np.pi * 2.0 * y_cut * pw * a[:, 0] * nev))
        a = col("dN_dphipdy", "dN_dphipdy"); out["dN_dphip"].append(a[:, 1] * (2.0 * y_cut * phw * nev))
        a = col("dN_taudtaudy", "dN_taudtaudy"); out["dN_tau"].append(a[:, 1] * (a[:, 0] * tw * nev * 2.0 * y_cut))
        a = col("dN_2pirdrdy", "dN_2pirdrdy"); out["dN_r"].append(a[:, 1] * (2.0 * np.pi * a[:, 0] * rw * nev * 2.0 * y_cut))
        a = col("dN_dphisdy", "dN_dphisdy"); out["dN_phis"].append(a[:, 1] * (phw * nev * 2.0 * y_cut))
    res = {k: np.rint(np.array(v)) for k, v in out.items()}
    res["nevents"] = nev
    res["total_yield"] = float(np.fromfile(os.path.join(root, "ref_dump", "total_yield.bin"), dtype=np.float64)[0])
    return res
