"""Seeded synthetic freezeout surfaces (SURVEY.md §8d "S-3D", "S-bundled", "S-VAH") and text writers for the
reference's surface.dat formats.  Pure numpy; used by tests, bench.py and the golden-vector generator.

Column contracts follow the reference readers:
  mode 1 (CPU VH/VAH)  reference src/cpp/readindata.cpp:222-295
  mode 6 (MUSIC)       reference src/cpp/readindata.cpp:383-508
"""
from __future__ import annotations

import numpy as np

HBARC = 0.197327053  # GeV fm, reference src/cpp/iS3D.h:14

# order of the structure-of-arrays columns handed to is3d_set_surface (include/is3d_b200.h)
SOA_COLUMNS = ("tau", "x", "y", "eta", "dat", "dax", "day", "dan", "ux", "uy", "un", "E", "T", "P",
               "pixx", "pixy", "pixn", "piyy", "piyn", "bulkPi", "muB", "nB", "Vx", "Vy", "Vn")


def bundled_cell() -> dict:
    """The single static thermal cell shipped as input/surface.dat, re-expressed in physical units
    (V = 100 fm^3, T = 0.760295 fm^-1, E = 1.40186 fm^-4, P = 0.20914 fm^-4)."""
    z = np.zeros(1)
    s = {k: z.copy() for k in SOA_COLUMNS}
    s["tau"][:] = 1.0
    s["dat"][:] = 100.0
    s["E"][:] = 1.40186 * HBARC
    s["T"][:] = 0.760295 * HBARC
    s["P"][:] = 0.20914 * HBARC
    return s


def s3d(n: int, seed: int = 12345, baryon: bool = False, dimension: int = 3, vah: bool = False,
        stress: float = 0.0) -> dict:
    """S-3D(N, seed): random 3+1D viscous-hydro cells in physical (GeV, fm) units as the kernels consume them.

    vah=True gives the S-VAH variant: a large pi^{eta eta}-dominated pressure anisotropy, P_L/P_T in [0.3, 1].
    dimension=2 zeroes eta, dsigma_eta, u^eta, pi^{x eta}, pi^{y eta}, V^eta (boost-invariant cells).
    """
    rng = np.random.default_rng(seed)
    u = rng.uniform
    s = {}
    tau = u(1.0, 10.0, n)
    s["tau"] = tau
    s["x"] = u(-8.0, 8.0, n)
    s["y"] = u(-8.0, 8.0, n)
    s["eta"] = u(-3.0, 3.0, n)
    s["dat"] = tau * u(0.01, 0.1, n)
    s["dax"] = u(-0.02, 0.02, n)
    s["day"] = u(-0.02, 0.02, n)
    s["dan"] = u(-0.02, 0.02, n)
    s["ux"] = u(-0.8, 0.8, n)
    s["uy"] = u(-0.8, 0.8, n)
    s["un"] = u(-0.1, 0.1, n) / tau
    T = u(0.140, 0.160, n)
    s["T"] = T
    s["E"] = 0.24 * (T / 0.15) ** 4
    P = 0.041 * (T / 0.15) ** 4
    s["P"] = P
    s["pixx"] = u(-0.1, 0.1, n) * P
    s["pixy"] = u(-0.1, 0.1, n) * P
    s["pixn"] = u(-0.1, 0.1, n) * P / tau
    s["piyy"] = u(-0.1, 0.1, n) * P
    s["piyn"] = u(-0.1, 0.1, n) * P / tau
    s["bulkPi"] = -u(0.0, 0.1, n) * P
    if vah:
        # pi^{zz}_LRF ~ -(pi^xx + pi^yy) for slow flow: push it negative so that P_L < P_T
        a = u(0.0, 0.35, n) * P
        s["pixx"] = 0.5 * a + u(-0.03, 0.03, n) * P
        s["piyy"] = 0.5 * a + u(-0.03, 0.03, n) * P
        s["pixy"] = u(-0.03, 0.03, n) * P
        s["pixn"] = u(-0.03, 0.03, n) * P / tau
        s["piyn"] = u(-0.03, 0.03, n) * P / tau
        s["ux"] = u(-0.4, 0.4, n)
        s["uy"] = u(-0.4, 0.4, n)
    if stress > 0.0:
        # a fraction `stress` of the cells gets large viscous corrections (|pi| ~ P, Pi ~ -0.6 P): these are the cells
        # where the modified distributions break down (detA <= deta_min, negative pion density, pl < 0)
        hot = u(0.0, 1.0, n) < stress
        for k in ("pixx", "pixy", "piyy"):
            s[k] = np.where(hot, 12.0 * s[k], s[k])
        for k in ("pixn", "piyn"):
            s[k] = np.where(hot, 12.0 * s[k], s[k])
        s["bulkPi"] = np.where(hot, -u(0.3, 0.9, n) * P, s["bulkPi"])
    if baryon:
        s["muB"] = u(0.05, 0.4, n)
        s["nB"] = u(0.01, 0.1, n)
        s["Vx"] = u(-1e-3, 1e-3, n)
        s["Vy"] = u(-1e-3, 1e-3, n)
        s["Vn"] = u(-1e-3, 1e-3, n) / tau
    else:
        for k in ("muB", "nB", "Vx", "Vy", "Vn"):
            s[k] = np.zeros(n)
    if dimension == 2:
        for k in ("eta", "dan", "un", "pixn", "piyn", "Vn"):
            s[k] = np.zeros(n)
    return {k: np.ascontiguousarray(s[k], dtype=np.float64) for k in SOA_COLUMNS}


BENCH_BLOCK = 1_250_000      # cells per block of the benchmark surface
BENCH_SEED = 2024


def bench_surface(begin: int, end: int, baryon: bool = True) -> dict:
    """Cells [begin, end) of THE benchmark surface (BASELINE.json config 5: synthetic 3+1D surface, 10 M cells): the
    concatenation of blocks of BENCH_BLOCK cells, block k = S-3D(BENCH_BLOCK, seed = BENCH_SEED + k).  The surface does not
    depend on how many GPUs share it; a rank generates only the blocks its cell range touches."""
    if not 0 <= begin <= end:
        raise ValueError(f"bad cell range [{begin}, {end})")
    parts = []
    for k in range(begin // BENCH_BLOCK, (max(end, 1) - 1) // BENCH_BLOCK + 1):
        blk = s3d(BENCH_BLOCK, seed=BENCH_SEED + k, baryon=baryon)
        lo, hi = max(begin - k * BENCH_BLOCK, 0), min(end - k * BENCH_BLOCK, BENCH_BLOCK)
        parts.append({c: v[lo:hi] for c, v in blk.items()})
    if len(parts) == 1:
        return {c: np.ascontiguousarray(v) for c, v in parts[0].items()}
    return {c: np.concatenate([p[c] for p in parts]) for c in SOA_COLUMNS}


def write_mode1(path: str, s: dict, baryon: bool = False) -> None:
    """surface.dat in the CPU-VH layout `t x y n ds_t ds_x ds_y ds_n u^x u^y u^n E T P pi^xx pi^xy pi^xn pi^yy
    pi^yn Pi [muB nB V^x V^y V^n]`, thermodynamic columns in fm^-1 units (reader multiplies by hbarc).
    17 significant digits; the value the reader reconstructs is (x / hbarc) * hbarc, which
    `roundtrip_mode1` reproduces so that both sides see bit-identical inputs."""
    cols = [s["tau"], s["x"], s["y"], s["eta"], s["dat"], s["dax"], s["day"], s["dan"], s["ux"], s["uy"], s["un"],
            s["E"] / HBARC, s["T"] / HBARC, s["P"] / HBARC, s["pixx"] / HBARC, s["pixy"] / HBARC, s["pixn"] / HBARC,
            s["piyy"] / HBARC, s["piyn"] / HBARC, s["bulkPi"] / HBARC]
    if baryon:
        cols += [s["muB"] / HBARC, s["nB"], s["Vx"], s["Vy"], s["Vn"]]
    np.savetxt(path, np.column_stack(cols), fmt="%.17e")


def roundtrip_mode1(s: dict, baryon: bool = False) -> dict:
    """The surface exactly as the mode-1 reader reconstructs it from `write_mode1` output."""
    out = {k: v.copy() for k, v in s.items()}
    keys = ["E", "T", "P", "pixx", "pixy", "pixn", "piyy", "piyn", "bulkPi"] + (["muB"] if baryon else [])
    for k in keys:
        out[k] = (s[k] / HBARC) * HBARC
    if not baryon:
        for k in ("muB", "nB", "Vx", "Vy", "Vn"):
            out[k] = np.zeros_like(s["tau"])
    return out


def write_mode6(path: str, s: dict, baryon: bool = False) -> None:
    """surface.dat in the MUSIC layout (reference src/cpp/readindata.cpp:383-508):
    `t x y n ds_t/t ds_x/t ds_y/t ds_n/t u^t u^x u^y t.u^n E T muB muS muC (E+P)/T pi^tt pi^tx pi^ty t.pi^tn
    pi^xx pi^xy t.pi^xn pi^yy t.pi^yn t2.pi^nn Pi [nB V^t V^x V^y t.V^n]`, energies in fm^-1 units."""
    tau = s["tau"]
    ut = np.sqrt(1.0 + s["ux"] ** 2 + s["uy"] ** 2 + (tau * s["un"]) ** 2)
    z = np.zeros_like(tau)
    E, T, P = s["E"] / HBARC, s["T"] / HBARC, s["P"] / HBARC
    cols = [tau, s["x"], s["y"], s["eta"], s["dat"] / tau, s["dax"] / tau, s["day"] / tau, s["dan"] / tau,
            ut, s["ux"], s["uy"], tau * s["un"], E, T, s["muB"] / HBARC, z, z, (E + P) / T,
            z, z, z, z, s["pixx"] / HBARC, s["pixy"] / HBARC, tau * s["pixn"] / HBARC, s["piyy"] / HBARC,
            tau * s["piyn"] / HBARC, z, s["bulkPi"] / HBARC]
    if baryon:
        cols += [s["nB"], z, s["Vx"], s["Vy"], tau * s["Vn"]]
    np.savetxt(path, np.column_stack(cols), fmt="%.17e")


def write_mode5(path: str, s: dict, baryon: bool = False, seed: int = 0) -> np.ndarray:
    """mode 1 columns followed by the six thermal-vorticity components wtx wty wtn wxy wxn wyn (reference
    src/cpp/readindata.cpp:299-307).  Returns the vorticity block that was written."""
    cols = [s["tau"], s["x"], s["y"], s["eta"], s["dat"], s["dax"], s["day"], s["dan"], s["ux"], s["uy"], s["un"],
            s["E"] / HBARC, s["T"] / HBARC, s["P"] / HBARC, s["pixx"] / HBARC, s["pixy"] / HBARC, s["pixn"] / HBARC,
            s["piyy"] / HBARC, s["piyn"] / HBARC, s["bulkPi"] / HBARC]
    if baryon:
        cols += [s["muB"] / HBARC, s["nB"], s["Vx"], s["Vy"], s["Vn"]]
    w = np.random.default_rng(seed).uniform(-0.05, 0.05, (len(s["tau"]), 6))
    np.savetxt(path, np.column_stack(cols + [w[:, k] for k in range(6)]), fmt="%.17e")
    return w


def write_mode7(path: str, s: dict) -> None:
    """surface.dat in the HIC-EventGen layout (2+1d, GeV units; reference src/cpp/readindata.cpp:589-690):
    `t x y n ds_t/t ds_x/t ds_y/t ds_n/t v^x v^y t.v^n pi^tt pi^tx pi^ty t.pi^tn pi^xx pi^xy t.pi^xn pi^yy t.pi^yn
    t2.pi^nn Pi T E P muB`."""
    tau = s["tau"]
    ut = np.sqrt(1.0 + s["ux"] ** 2 + s["uy"] ** 2)
    z = np.zeros_like(tau)
    cols = [tau, s["x"], s["y"], z, s["dat"] / tau, s["dax"] / tau, s["day"] / tau, z, s["ux"] / ut, s["uy"] / ut, z,
            z, z, z, z, s["pixx"], s["pixy"], z, s["piyy"], z, z, s["bulkPi"], s["T"], s["E"], s["P"], z]
    np.savetxt(path, np.column_stack(cols), fmt="%.17e")
