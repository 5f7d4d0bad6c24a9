"""Seeded sweep of parameter combinations: CUDA path (through the C ABI) against the CPU oracle on the same inputs.

The golden vectors of the unmodified reference pin the oracle on 34 spectra and 11 dN/dX cases (tests/test_oracle_cpu.py); this
sweep walks the switches those cases do not combine -- df_mode x dimension x {bulk, shear, baryon, baryon diffusion} x
regulate_deltaf x outflow x species list (pi/K/p or all SMASH species incl. antibaryons and the deuteron) -- so that every
template instantiation of the spectra and dN/dX kernels (MODE, BARYON, REGULATE, OUTFLOW, SPECIES_RENORM, LINEAR) is
evaluated at least once against the restated reference algorithm.  Tolerance: tests/harness.py (1e-10 relative per bin)."""
import itertools

import numpy as np
import pytest

import cases
import harness
import oracle_api
from is3d2_b200 import synthetic, workdir

pytestmark = pytest.mark.gpu


def _combos():
    rng = np.random.default_rng(20240919)
    out = []
    k = 0
    for df_mode, dim, baryon in itertools.product((1, 2, 3, 4, 5), (3, 2), (0, 1)):
        if df_mode == 4 and baryon:            # PTB has no muB != 0 coefficient tables (reference DeltafData.cpp:480-484)
            continue
        for rep in range(2):
            k += 1
            smash = (rep == 1)
            p = dict(df_mode=df_mode, dimension=dim, include_baryon=baryon,
                     include_baryondiff_deltaf=int(baryon and rng.integers(0, 2)),
                     include_bulk_deltaf=int(rng.integers(0, 4) > 0), include_shear_deltaf=int(rng.integers(0, 4) > 0),
                     regulate_deltaf=int(rng.integers(0, 2)), outflow=int(rng.integers(0, 2)),
                     hrg_eos=2 if smash else int(rng.integers(1, 3)))
            n = int(rng.integers(5, 9)) if smash else int(rng.integers(40, 300))
            out.append((f"m{df_mode}_{dim}d_b{baryon}_{'smash' if smash else 'pikp'}_{k}", p, n, smash, 1000 + k))
    return out


COMBOS = _combos()


def _surface(p, n, seed):
    stress = 0.3 if p["df_mode"] in (3, 4) else 0.0
    return synthetic.roundtrip_mode1(synthetic.s3d(n, seed=seed, baryon=bool(p["include_baryon"]), dimension=p["dimension"],
                                                   stress=stress, vah=(p["df_mode"] == 5)), baryon=bool(p["include_baryon"]))


@pytest.mark.parametrize("name,p,n,smash,seed", COMBOS, ids=[c[0] for c in COMBOS])
def test_spectra_sweep_matches_oracle(libs, tmp_path, name, p, n, smash, seed):
    params = cases._p(**p)
    case = dict(params=params, chosen="smash" if smash else "pikp")
    surf = _surface(p, n, seed)
    with harness.open_session(str(tmp_path / "gpu"), case, surf) as h:
        got, st = h.abi_spectra()
    root = workdir.make_workdir(str(tmp_path / "oracle"), params, chosen=case["chosen"])
    rc, want, ost = oracle_api.OracleProblem(root, params, surf).spectra()
    assert rc == 0
    worst = harness.assert_spectra_close(got, want, what=name)
    assert st.cells_skipped == ost.cells_skipped and st.cells_breakdown == ost.cells_breakdown
    print(f"{name}: {n} cells, max rel err {worst:.2e}")


DNDX_COMBOS = [c for c in COMBOS if c[1]["df_mode"] != 5]


@pytest.mark.parametrize("name,p,n,smash,seed", DNDX_COMBOS, ids=[c[0] for c in DNDX_COMBOS])
def test_dndx_sweep_matches_oracle(libs, tmp_path, name, p, n, smash, seed):
    import ctypes as C
    from is3d2_b200 import Stats
    params = cases._p(operation=0, **p)
    case = dict(params=params, chosen="smash" if smash else "pikp")
    n = min(n, 60)
    surf = _surface(p, n, seed + 500)
    root = workdir.make_workdir(str(tmp_path / "oracle"), params, chosen=case["chosen"])
    rc, want, _ = oracle_api.OracleProblem(root, params, surf).dndx()
    assert rc == 0
    ns = want["tau"].shape[0]
    got = {"tau": np.zeros_like(want["tau"]), "r": np.zeros_like(want["r"]), "phi": np.zeros_like(want["phi"])}
    with harness.open_session(str(tmp_path / "gpu"), case, surf) as h:
        st = Stats()
        rc = h.lib.is3d_dndx(h.ctx, got["tau"].ctypes.data, got["r"].ctypes.data, got["phi"].ctypes.data, C.byref(st))
        assert rc == 0, h.lib.is3d_last_error(h.ctx)
    for k in ("tau", "r", "phi"):
        assert got[k].shape == (ns, want[k].shape[1])
        harness.assert_hist_close(got[k], want[k], what=f"{name}/{k}")


@pytest.mark.parametrize("df_mode,operation", [(2, 1), (3, 1), (1, 0)])
def test_urqmd_list_with_baryon_terms_matches_oracle(libs, tmp_path, df_mode, operation):
    """All 305 UrQMD species with baryon terms: another particle list for the class / charge-conjugate pair layout (its baryon
    and antibaryon multiplets pair up differently from SMASH's), spectra and dN/dX against the CPU oracle."""
    import ctypes as C
    from is3d2_b200 import Stats
    p = dict(df_mode=df_mode, dimension=3, include_baryon=1, include_baryondiff_deltaf=1, hrg_eos=1, operation=operation)
    params = cases._p(**p)
    case = dict(params=params, chosen="urqmd_v3.3+")
    surf = synthetic.roundtrip_mode1(synthetic.s3d(7, seed=4242 + df_mode, baryon=True, stress=0.3 if df_mode == 3 else 0.0), baryon=True)
    root = workdir.make_workdir(str(tmp_path / "oracle"), params, chosen=case["chosen"])
    prob = oracle_api.OracleProblem(root, params, surf)
    with harness.open_session(str(tmp_path / "gpu"), case, surf) as h:
        if operation == 1:
            got, st = h.abi_spectra()
            rc, want, _ = prob.spectra()
            assert rc == 0 and got.shape[0] == 305
            worst = harness.assert_spectra_close(got, want, what=f"urqmd df_mode {df_mode}")
            assert st.pair_evals_executed > 0 or df_mode > 2          # K1 reports its pair slots
        else:
            rc, want, _ = prob.dndx()
            assert rc == 0
            got = {k: np.zeros_like(v) for k, v in want.items()}
            stt = Stats()
            h._check(h.lib.is3d_dndx(h.ctx, got["tau"].ctypes.data_as(C.c_void_p), got["r"].ctypes.data_as(C.c_void_p),
                                     got["phi"].ctypes.data_as(C.c_void_p), C.byref(stt)), "is3d_dndx")
            for k in ("tau", "r", "phi"):
                harness.assert_hist_close(got[k], want[k], what=f"urqmd dN/dX {k}")
