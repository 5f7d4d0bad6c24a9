"""GPU parity of the spacetime distributions dN/dX (K4, df_mode 1-4) against the reference's golden files."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import harness

pytestmark = pytest.mark.gpu


def _device_hists(h, ns, params):
    tb, rb, pb = (int(float(params.get(k, d))) for k, d in (("tau_bins", 120), ("r_bins", 60), ("phip_bins", 100)))
    tau, r, phi = np.zeros((ns, tb)), np.zeros((ns, rb)), np.zeros((ns, pb))
    from is3d2_b200 import Stats
    st = Stats()
    rc = h.lib.is3d_dndx(h.ctx, tau.ctypes.data, r.ctypes.data, phi.ctypes.data, C.byref(st))
    assert rc == 0, h.lib.is3d_last_error(h.ctx)
    return {"tau": tau, "r": r, "phi": phi}, st


@pytest.mark.parametrize("name", list(cases.DNDX_CASES) + list(cases.BIG_DNDX_CASES))
def test_dndx_matches_reference(libs, tmp_path, name):
    case = cases.DNDX_CASES.get(name) or cases.BIG_DNDX_CASES[name]
    surf, ref = harness.load_golden_dndx(name)
    ns = ref["tau"].shape[0]
    with harness.open_session(str(tmp_path), case, surf) as h:
        clean, st = _device_hists(h, ns, case["params"])
        assert st.cells_total == len(surf["tau"])
    got = harness.normalise_dndx({k: harness.emulate_partial_memset(v) for k, v in clean.items()}, case["params"])
    for k in ("tau", "r", "phi"):
        harness.assert_hist_close(got[k], ref[k], what=f"{name}/{k}")


@pytest.mark.parametrize("name", ["dndx_bench_m2_smash_baryon_512cells", "dndx_s3d_m3"])
def test_dndx_negligible_margin_is_checked_a_posteriori(libs, tmp_path, monkeypatch, name):
    """is3d_params.negligible_margin in K4 (df and feqmod kernels): quadrature points far above the cell's smallest exponent are
    not marched over; their summed bound is tested against every (cell, class) scalar.  Default margin == margin off to 1e-12 per
    bin (FP64 atomics: no bitwise claim), points are dropped, no rerun; an absurd margin trips the test and the rerun delivers
    the margin-off histograms."""
    case = cases.DNDX_CASES.get(name) or cases.BIG_DNDX_CASES[name]
    surf, ref = harness.load_golden_dndx(name)
    ns = ref["tau"].shape[0]
    out = {}
    for tag, margin in (("off", "0"), ("default", None), ("absurd", "1e-3")):
        if margin is None:
            monkeypatch.delenv("IS3D_NEGLIGIBLE_MARGIN", raising=False)
        else:
            monkeypatch.setenv("IS3D_NEGLIGIBLE_MARGIN", margin)
        with harness.open_session(str(tmp_path / tag), case, surf) as h:
            out[tag] = _device_hists(h, ns, case["params"])
    (off, st_off), (dflt, st_d), (absurd, st_a) = out["off"], out["default"], out["absurd"]
    assert st_off.prune_reruns == 0 and st_d.prune_reruns == 0 and st_a.prune_reruns == 1
    assert st_d.evals_dropped > st_off.evals_dropped
    for k in ("tau", "r", "phi"):
        scale = np.abs(off[k]).max()
        assert np.abs(dflt[k] - off[k]).max() <= 1e-12 * scale
        assert np.abs(absurd[k] - off[k]).max() <= 1e-12 * scale
    print(f"{name}: thread-slot evaluations dropped {st_d.evals_dropped:.3g} (margin off {st_off.evals_dropped:.3g})")


def test_dndx_host_bug_compat_and_files(libs, tmp_path):
    """Through the host layer (EmissionFunctionArray::calculate_spectra, operation 0): the bug-compatible mode
    reproduces the reference's files; the default writes clean per-species histograms whose total over bins equals
    the pT/phi/y-integrated yield."""
    name = "dndx_s3d_m1"
    case = cases.DNDX_CASES[name]
    surf, ref = harness.load_golden_dndx(name)
    os.environ["IS3D_DNDX_BUG_COMPAT"] = "1"
    try:
        with harness.open_session(str(tmp_path / "bug"), case, surf) as h:
            h.run()
            tau = C.POINTER(C.c_double)(); r = C.POINTER(C.c_double)(); phi = C.POINTER(C.c_double)()
            ns = h.host.is3d_host_dndx(h.h, C.byref(tau), C.byref(r), C.byref(phi))
            got = {"tau": np.ctypeslib.as_array(tau, shape=(ns, 120)).copy(), "r": np.ctypeslib.as_array(r, shape=(ns, 60)).copy(),
                   "phi": np.ctypeslib.as_array(phi, shape=(ns, 100)).copy()}
    finally:
        del os.environ["IS3D_DNDX_BUG_COMPAT"]
    got = harness.normalise_dndx(got, case["params"])
    for k in ("tau", "r", "phi"):
        harness.assert_hist_close(got[k], ref[k], what=f"host/{k}")
    # file layout: bin mid-point <tab> value, %.6e, one file per species and histogram
    f = tmp_path / "bug" / "results" / "continuous" / "dN_taudtaudy_211.dat"
    rows = np.loadtxt(f)
    assert rows.shape == (120, 2)
    np.testing.assert_allclose(rows[:, 1], ref["tau"][0], rtol=2e-6, atol=1e-300)
