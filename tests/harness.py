"""Shared helpers of the parity tests."""
from __future__ import annotations

import os

import numpy as np

import cases
from is3d_b200 import HostSession, synthetic, workdir

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated tolerance of the continuous paths (BASELINE.json north_star): 1e-10 relative per bin in FP64.
# Bins more than 200 decades below the largest one are compared absolutely (the reference underflows to 0 there).
RTOL = 1e-10


def load_golden(name: str):
    z = np.load(os.path.join(GOLDEN, f"spectra_{name}.npz"))
    surf = {k[4:]: z[k] for k in z.files if k.startswith("col_")}
    return surf, z["spectra"]


def assert_spectra_close(got: np.ndarray, ref: np.ndarray, rtol: float = RTOL, what: str = ""):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.all(np.isfinite(got)), f"{what}: non-finite values"
    scale = np.abs(ref).max()
    floor = scale * 1e-200
    err = np.abs(got - ref)
    bad = err > rtol * np.abs(ref) + floor
    if bad.any():
        i = np.unravel_index(np.argmax(err / (np.abs(ref) + floor)), ref.shape)
        raise AssertionError(f"{what}: {bad.sum()} of {ref.size} bins differ by more than {rtol:g} relative; worst at {i}: "
                             f"got {got[i]!r} ref {ref[i]!r}")
    big = np.abs(ref) > floor
    return float((err[big] / np.abs(ref[big])).max()) if big.any() else 0.0


def open_session(tmp: str, case: dict, surface: dict, overrides: dict | None = None) -> HostSession:
    """Working directory + host session + in-memory surface + CUDA context for one parity case."""
    params = dict(case["params"])
    if overrides:
        params.update(overrides)
    root = workdir.make_workdir(tmp, params, chosen=case["chosen"], **case.get("tables", {}))
    h = HostSession(root)
    h.set_surface(surface)
    h.prepare()
    return h
