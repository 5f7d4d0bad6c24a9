"""Shared helpers of the parity tests."""
from __future__ import annotations

import os

import numpy as np

import cases
from is3d2_b200 import HostSession, synthetic, workdir

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated tolerance of the continuous paths (BASELINE.json north_star): 1e-10 relative per bin in FP64.
# Two absolute floors: bins more than 200 decades below the largest one (the reference underflows to 0 there), and
# 1e-13 of the species' largest bin -- a few bins are sums of positive and negative p.dsigma contributions that
# cancel to 1e-4..1e-6 of their gross size (even to a negative net value); any two correct FP64 summations differ
# there by (gross size) x 1e-16, which exceeds 1e-10 of the NET value.  1e-13 of the species peak keeps those bins
# honest without asking for more digits than the reference itself carries.
RTOL = 1e-10
ATOL_OF_SPECIES_PEAK = 1e-13


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN, f"spectra_{name}.npz"))
    surf = {k[4:]: z[k] for k in z.files if k.startswith("col_")}
    return surf, z["spectra"]


def assert_spectra_close(got: np.ndarray, ref: np.ndarray, rtol: float = RTOL, what: str = ""):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.all(np.isfinite(got)), f"{what}: non-finite values"
    scale = np.abs(ref).max()
    peak = np.abs(ref).reshape(ref.shape[0], -1).max(axis=1).reshape((-1,) + (1,) * (ref.ndim - 1))
    floor = scale * 1e-200 + ATOL_OF_SPECIES_PEAK * peak
    err = np.abs(got - ref)
    bad = err > rtol * np.abs(ref) + floor
    if bad.any():
        i = np.unravel_index(np.argmax(err / (np.abs(ref) + floor)), ref.shape)
        raise AssertionError(f"{what}: {bad.sum()} of {ref.size} bins differ by more than {rtol:g} relative; worst at {i}: "
                             f"got {got[i]!r} ref {ref[i]!r}")
    big = np.abs(ref) > 1e3 * floor
    return float((err[big] / np.abs(ref[big])).max()) if big.any() else 0.0


def open_session(tmp: str, case: dict, surface: dict, overrides: dict | None = None, famod_chain: int = 1) -> HostSession:
    """Working directory + host session + in-memory surface + CUDA context for one parity case.  famod_chain = 1 (df_mode 5
    only): the reference's serial initial-guess chain, which the multi-cell goldens of the unmodified reference carry; the
    library default is 0 (chain-free), tested against the one-cell reference runs (load_golden_m5free)."""
    params = dict(case["params"])
    if overrides:
        params.update(overrides)
    root = workdir.make_workdir(tmp, params, chosen=case["chosen"], **case.get("tables", {}))
    h = HostSession(root)
    h.set_surface(surface)
    old = os.environ.get("IS3D_FAMOD_CHAIN")
    os.environ["IS3D_FAMOD_CHAIN"] = str(famod_chain)          # read once, by the EmissionFunctionArray constructor
    try:
        h.prepare()
    finally:
        if old is None:
            os.environ.pop("IS3D_FAMOD_CHAIN", None)
        else:
            os.environ["IS3D_FAMOD_CHAIN"] = old
    return h


# ---- dN/dX helpers -----------------------------------------------------------------------------------------------
def load_golden_dndx(name: str):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    surf = {k[4:]: z[k] for k in z.files if k.startswith("col_")}
    return surf, {"tau": z["tau"], "r": z["r"], "phi": z["phi"]}


def emulate_partial_memset(h: np.ndarray) -> np.ndarray:
    """The reference reuses one accumulator per histogram for all species and clears it with
    memset(ptr, 0.0, CORES * bins) -- `bins` BYTES (SpacetimeDistribution.cpp:166-168): serial build, literally."""
    ns, bins = h.shape
    acc = np.zeros(bins)
    out = np.empty_like(h)
    for s in range(ns):
        acc.view(np.uint8)[:bins] = 0
        acc += h[s]
        out[s] = acc
    return out


def normalise_dndx(h: dict, params: dict) -> dict:
    """Writer normalisation of SpacetimeDistribution.cpp:448-490."""
    from is3d2_b200 import workdir
    p = workdir.default_parameters()
    p.update({k: str(v) for k, v in params.items()})
    tb, rb, pb = int(float(p["tau_bins"])), int(float(p["r_bins"])), int(float(p["phip_bins"]))
    tw = (float(p["tau_max"]) - float(p["tau_min"])) / tb
    rw = (float(p["r_max"]) - float(p["r_min"])) / rb
    pw = 2.0 * np.pi / pb
    tau_mid = float(p["tau_min"]) + tw * (np.arange(tb) + 0.5)
    r_mid = float(p["r_min"]) + rw * (np.arange(rb) + 0.5)
    return {"tau": h["tau"] / (tau_mid * tw), "r": h["r"] / (2.0 * np.pi * r_mid * rw), "phi": h["phi"] / pw}


def assert_hist_close(got: np.ndarray, ref: np.ndarray, rtol: float = RTOL, what: str = ""):
    assert got.shape == ref.shape
    peak = np.abs(ref).max(axis=1, keepdims=True)
    err = np.abs(got - ref)
    bad = err > rtol * np.abs(ref) + ATOL_OF_SPECIES_PEAK * peak
    assert not bad.any(), f"{what}: {bad.sum()} bins off; worst {np.max(err / (np.abs(ref) + 1e-300 + ATOL_OF_SPECIES_PEAK * peak)):.3e}"


# ---- sampler helpers ---------------------------------------------------------------------------------------------
def load_golden_sampler(name: str):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    surf = {k[4:]: z[k] for k in z.files if k.startswith("col_")}
    ref = {k: z[k] for k in z.files if not k.startswith("col_")}
    return surf, ref


def chi2_two_sample(a, sa, b, sb, min_counts=100):
    """Two-sample chi^2 of count histograms a (sa events) and b (sb events) with Poisson variances, over bins with at
    least `min_counts` entries in total (below that the observed-variance estimate is biased)."""
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    m = (a + b) >= min_counts
    d = a[m] / sa - b[m] / sb
    v = a[m] / sa ** 2 + b[m] / sb ** 2
    return float((d * d / v).sum()), int(m.sum())


# ---- spin polarization helpers -----------------------------------------------------------------------------------
def load_golden_polzn(name: str):
    """(surface columns, vorticity (6, n), reference arrays (5, Ns, NpT, Nphi, Ny)); large cases regenerate their inputs
    from the seeds exactly as tests/golden/make_golden_polzn.py wrote them."""
    import tempfile
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    if "vorticity" in z.files:
        surf = {k[4:]: z[k] for k in z.files if k.startswith("col_")}
        return surf, z["vorticity"], z["polarization"]
    case = cases.POLZN_CASES[name]
    s = cases.make_surface(case["surface"])
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "surface.dat")
        synthetic.write_mode5(p, s, baryon=False, seed=len(s["tau"]))
        flat = np.loadtxt(p, ndmin=2)
    return synthetic.roundtrip_mode1(s, baryon=False), np.ascontiguousarray(flat[:, 20:26].T), z["polarization"]


def assert_polzn_close(got: np.ndarray, ref: np.ndarray, rtol: float = RTOL, what: str = ""):
    """Per component and species: |got - ref| <= rtol |ref| + 1e-13 of the component's largest bin of that species."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.all(np.isfinite(got)), f"{what}: non-finite values"
    peak = np.abs(ref).reshape(ref.shape[0], ref.shape[1], -1).max(axis=2).reshape(ref.shape[:2] + (1, 1, 1))
    err = np.abs(got - ref)
    bad = err > rtol * np.abs(ref) + ATOL_OF_SPECIES_PEAK * peak
    if bad.any():
        i = np.unravel_index(np.argmax(err / (np.abs(ref) + ATOL_OF_SPECIES_PEAK * peak + 1e-300)), ref.shape)
        raise AssertionError(f"{what}: {bad.sum()} of {ref.size} entries differ by more than {rtol:g}; worst at {i}: got {got[i]!r} ref {ref[i]!r}")
    big = np.abs(ref) > 1e3 * ATOL_OF_SPECIES_PEAK * peak
    return float((err[big] / np.abs(ref[big])).max()) if big.any() else 0.0


def load_golden_m5free(name: str):
    """df_mode 5 chain-free golden: sum of one-cell runs of the unmodified reference (make_golden_m5_chainfree.py)."""
    z = np.load(os.path.join(GOLDEN, f"m5free_{name}.npz"))
    surf = {k[4:]: z[k] for k in z.files if k.startswith("col_")}
    return surf, z["spectra"]
