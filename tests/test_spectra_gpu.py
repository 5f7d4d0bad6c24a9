"""GPU parity of the continuous Cooper-Frye spectra (K1: df_mode 1, 2; K2: df_mode 3, 4; K3: df_mode 5) through the C ABI against the golden
vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

import cases
import harness

pytestmark = pytest.mark.gpu

DF_CASES = [n for n, c in cases.SPECTRA_CASES.items() if c["params"]["df_mode"] in (1, 2, 3, 4, 5)]


@pytest.mark.parametrize("name", DF_CASES)
def test_spectra_df_matches_reference(libs, tmp_path, name):
    case = cases.SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    with harness.open_session(str(tmp_path), case, surf) as h:
        got, st = h.abi_spectra()
        worst = harness.assert_spectra_close(got, ref, what=name)
        assert st.cells_total == len(surf["tau"])
        assert st.kernel_launches >= 3
        # same call again: deterministic reduction => bit-identical
        again, _ = h.abi_spectra()
        np.testing.assert_array_equal(got, again)
    print(f"{name}: max rel err {worst:.3e}")


M5_CASES = [n for n, c in cases.SPECTRA_CASES.items() if c["params"]["df_mode"] == 5 and c["chosen"] == "pikp"]


@pytest.mark.parametrize("name", M5_CASES)
def test_famod_chain_free_fast_path_matches_oracle(libs, tmp_path, monkeypatch, name):
    """df_mode 5 production mode (famod_chain = 0): every cell's Newton solve starts from (T, 1, 1) and the term sums use
    the FP64-pipe approximations (fast_exp / fast_atan / rcp / rsqrt).  Checked against the CPU oracle run with the same
    chain-free policy.  The closed-form angular functions cancel by up to 3e4 near z = 0.01, so two correct evaluations
    with different 1-ulp roundings of atan differ at the 1e-11 level; the tolerance here is 5e-10 and the observed worst
    case is printed (profiles/ records it)."""
    import oracle_api
    from is3d2_b200 import workdir
    case = cases.SPECTRA_CASES[name]
    surf, _ = harness.load_golden(name)
    with harness.open_session(str(tmp_path / "gpu"), case, surf, famod_chain=0) as h:
        got, st = h.abi_spectra()
    root = workdir.make_workdir(str(tmp_path / "oracle"), case["params"], chosen=case["chosen"], **case.get("tables", {}))
    rc, ref, ost = oracle_api.OracleProblem(root, case["params"], surf, famod_chain=0).spectra()
    assert rc == 0
    worst = harness.assert_spectra_close(got, ref, rtol=5e-10, what=name + " chain-free")
    assert st.newton_iterations == ost.newton_iterations
    assert st.cells_breakdown == ost.cells_breakdown and st.reconstruction_failures == ost.reconstruction_failures
    print(f"{name} chain-free fast path vs oracle: max rel err {worst:.3e}")


@pytest.mark.parametrize("name", list(cases.M5_CHAINFREE_CASES))
def test_famod_chain_free_matches_one_cell_reference_runs(libs, tmp_path, name):
    """df_mode 5 in the library's default (production, shardable) policy -- every cell's Newton solve starts from (T, 1, 1) --
    against the UNMODIFIED REFERENCE: the golden is the sum of one-cell reference runs, where the reference itself starts
    from (T, 1, 1) because a one-cell surface has no previous solution (MomentumSpectra.cpp:1288-1313;
    tests/golden/make_golden_m5_chainfree.py).  Tolerance: the north-star 1e-10."""
    case = cases.M5_CHAINFREE_CASES[name]
    surf, ref = harness.load_golden_m5free(name)
    with harness.open_session(str(tmp_path), case, surf, famod_chain=0) as h:
        got, st = h.abi_spectra()
    worst = harness.assert_spectra_close(got, ref, what=name + " chain-free vs one-cell reference runs")
    assert st.cells_total == len(surf["tau"])
    print(f"{name} chain-free vs reference: max rel err {worst:.3e}")


@pytest.mark.parametrize("name", list(cases.BIG_SPECTRA_CASES))
def test_launch_realistic_surface_matches_reference(libs, tmp_path, name):
    """BASELINE.json config 2 on a prefix of THE benchmark surface (config 5) large enough for the launch shape bench.py
    times -- several cell chunks x 10 column slices of 256 thread columns x 21 rapidity blocks, all 444 SMASH species in 193
    classes / 50 uniform-baryon thread groups -- against one serial run of the unmodified reference
    (MomentumSpectra.cpp:99-375).  Also through small passes (multi-pass accumulation across pass boundaries)."""
    case = cases.BIG_SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    gen = cases.make_surface(case["surface"])          # the golden's cells ARE the benchmark surface's first cells
    for k in ("tau", "ux", "dat"):
        np.testing.assert_array_equal(gen[k], surf[k])
    with harness.open_session(str(tmp_path), case, surf) as h:
        got, st = h.abi_spectra()
    worst = harness.assert_spectra_close(got, ref, what=name)
    assert st.cells_total == len(surf["tau"]) and st.cells_out_of_table == 0
    print(f"{name}: max rel err {worst:.3e}, {st.cells_skipped} of {st.cells_total} cells skipped")


@pytest.mark.parametrize("name", list(cases.BIG_SPECTRA_CASES))
def test_negligible_margin_is_checked_a_posteriori(libs, tmp_path, monkeypatch, name):
    """is3d_params.negligible_margin (K1: df_mode 2 case, K2: df_mode 3 case): items far above a block row's smallest exponent
    are dropped before the momentum loop, the bounds of what was dropped are compared with every finished bin, and the call is
    repeated without the margin when a bin fails.  (a) default margin on the benchmark surface's first cells: fewer
    evaluations executed, no rerun, same spectra as with the margin off to 1e-13; (b) an absurd margin that drops leading
    terms: the test fails, the rerun delivers the margin-off result bit for bit."""
    case = cases.BIG_SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    out = {}
    for tag, margin in (("off", "0"), ("default", None), ("absurd", "1e-3")):
        if margin is None:
            monkeypatch.delenv("IS3D_NEGLIGIBLE_MARGIN", raising=False)
        else:
            monkeypatch.setenv("IS3D_NEGLIGIBLE_MARGIN", margin)
        with harness.open_session(str(tmp_path / tag), case, surf) as h:
            out[tag] = h.abi_spectra()
    (off, st_off), (dflt, st_d), (absurd, st_a) = out["off"], out["default"], out["absurd"]
    assert st_off.prune_reruns == 0 and st_d.prune_reruns == 0 and st_a.prune_reruns == 1
    assert st_d.evals_executed < 0.9 * st_off.evals_executed, (st_d.evals_executed, st_off.evals_executed)
    np.testing.assert_array_equal(absurd, off)
    big = np.abs(off) > 1e-250
    assert np.abs(dflt[big] / off[big] - 1.0).max() < 1e-13
    harness.assert_spectra_close(dflt, ref, what=name + " default margin")
    print(f"default margin executes {st_d.evals_executed / st_off.evals_executed:.3f} of the margin-off class evaluations")


def test_known_answer_static_cell(libs, tmp_path):
    """Ideal static cell: dN = g/(2 pi hbarc)^3 mT cosh(eta) dsigma_tau feq, and the reference's own
    printed value for pi+ at pT = 0 (1.37049908e+01)."""
    case = cases.SPECTRA_CASES["bundled_m1_2d"]
    surf, ref = harness.load_golden("bundled_m1_2d")
    with harness.open_session(str(tmp_path), case, surf) as h:
        got, _ = h.abi_spectra()
    assert abs(got[0, 0, 0, 0] / 1.37049908e+01 - 1) < 1e-8
    harness.assert_spectra_close(got, ref, what="bundled")


def test_linearity_in_cells(libs, tmp_path):
    """Size-independent property: spectra are additive over cells -- two halves sum to the whole, and a
    surface with every cell duplicated gives exactly twice the spectra (same summation tree per bin is not
    guaranteed, so compare at 1e-12)."""
    case = cases.SPECTRA_CASES["s3d_m2"]
    surf, ref = harness.load_golden("s3d_m2")
    n = len(surf["tau"])
    with harness.open_session(str(tmp_path), case, surf) as h:
        whole, _ = h.abi_spectra()
        a = {k: v[: n // 3] for k, v in surf.items()}
        b = {k: v[n // 3:] for k, v in surf.items()}
        h.abi_set_surface(a)
        sa, _ = h.abi_spectra()
        h.abi_set_surface(b)
        sb, _ = h.abi_spectra()
        h.abi_set_surface({k: np.concatenate([v, v]) for k, v in surf.items()})
        twice, _ = h.abi_spectra()
    harness.assert_spectra_close(sa + sb, whole, rtol=1e-11, what="halves")
    harness.assert_spectra_close(twice, 2.0 * whole, rtol=1e-11, what="duplicated")


def test_skipped_cells_and_empty_surface(libs, tmp_path):
    """Cells with u.dsigma <= 0 contribute nothing (reference MomentumSpectra.cpp:132); an empty surface gives zeros."""
    case = cases.SPECTRA_CASES["s3d_m1"]
    surf, ref = harness.load_golden("s3d_m1")
    flipped = {k: v.copy() for k, v in surf.items()}
    for k in ("dat", "dax", "day", "dan"):
        flipped[k] = -flipped[k]
    both = {k: np.concatenate([surf[k], flipped[k]]) for k in surf}
    with harness.open_session(str(tmp_path), case, surf) as h:
        h.abi_set_surface(both)
        got, st = h.abi_spectra()
        assert st.cells_skipped == len(surf["tau"])
        harness.assert_spectra_close(got, ref, what="skipped")
        h.abi_set_surface({k: v[:0] for k, v in surf.items()})
        zero, _ = h.abi_spectra()
        assert np.all(zero == 0.0)


def test_out_of_table_cell_is_an_error(libs, tmp_path):
    """T outside [0.1, 0.2] GeV aborts the reference (GSL domain error); the ABI returns IS3D_ERR_TABLE_RANGE."""
    from is3d2_b200 import Is3dError
    case = cases.SPECTRA_CASES["s3d_m1"]
    surf, _ = harness.load_golden("s3d_m1")
    bad = {k: v.copy() for k, v in surf.items()}
    bad["T"][5] = 0.25
    with harness.open_session(str(tmp_path), case, surf) as h:
        h.abi_set_surface(bad)
        with pytest.raises(Is3dError, match="status 3"):
            h.abi_spectra()


def test_host_writers_file_layout(libs, tmp_path):
    """operation 1 through the host layer (EmissionFunctionArray::calculate_spectra + the five writers, reference
    EmissionFunction.cpp:406-558, :804-878; layouts of SURVEY.md appendix C): every file is re-derived from the in-memory
    spectra -- header, loop order iy -> iphi -> ipT, blank-line structure, quadrature sums, 9 significant digits."""
    name = "s2d_m1_phi48"                                  # 48 phi points, 2+1d: y = 0 single row
    case = cases.SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    with harness.open_session(str(tmp_path), case, surf) as h:
        h.run()
        spec = h.spectra()                                 # (Ns, NpT, Nphi, Ny)
        mcid = h.chosen()
    harness.assert_spectra_close(spec, ref, what="host run")
    tab = lambda f: np.loadtxt(tmp_path / "tables" / "momentum" / f, ndmin=2)  # noqa: E731
    pT, phi = tab("pT_table.dat"), tab("phi_table.dat")
    ns, npT, nphi, ny = spec.shape
    assert (nphi, ny) == (48, 1)
    out = tmp_path / "results" / "continuous"
    for s, m in enumerate(mcid):
        lines = open(out / f"dN_pTdpTdphidy_{m}.dat").read().split("\n")
        assert lines[0] == "y\tphip\tpT\tdN_pTdpTdphidy"
        body = lines[1:]
        assert len(body) == ny * nphi * (npT + 1) + 1 and body[npT] == "" and body[-1] == ""     # blank line after each phi block
        rows = np.array([[float(v) for v in l.split("\t")] for l in body if l])
        want = np.transpose(spec[s], (2, 1, 0)).reshape(-1)                                      # iy -> iphi -> ipT
        np.testing.assert_allclose(rows[:, 3], want, rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(rows[:npT, 2], pT[:, 0], rtol=1e-8)
        np.testing.assert_allclose(rows[::npT, 1][:nphi], phi[:, 0], rtol=1e-8)
        d = np.loadtxt(out / f"dN_2pipTdpTdy_{m}.dat", ndmin=2)
        np.testing.assert_allclose(d[:, 2], (spec[s, :, :, 0] * phi[:, 1]).sum(axis=1) / (2 * np.pi), rtol=1e-8)
        d = np.loadtxt(out / f"dN_dphidy_{m}.dat", ndmin=2)
        np.testing.assert_allclose(d[:, 2], (spec[s, :, :, 0] * pT[:, 1][:, None]).sum(axis=0), rtol=1e-8)
        d = np.loadtxt(out / f"dN_dy_{m}.dat", ndmin=2)
        np.testing.assert_allclose(d[0, 1], (spec[s, :, :, 0] * pT[:, 1][:, None] * phi[:, 1][None, :]).sum(), rtol=1e-7)
        d = np.loadtxt(out / f"vn_{m}.dat", ndmin=2)
        assert d.shape == (npT, 9)
        w = spec[s, :, :, 0] * phi[:, 1]
        v2 = np.abs((w * np.exp(2j * phi[:, 0])).sum(axis=1)) / w.sum(axis=1)
        np.testing.assert_allclose(d[:, 3], v2, rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("name", ["s3d_m2_baryon", "s3d_m3", "s3d_m4", "vah_m5"])
def test_multi_pass_accumulation(libs, tmp_path, monkeypatch, name):
    """Surfaces beyond one pass (2-4 M cells) are streamed pass by pass into the same partial sums; forced here with the
    IS3D_PASS_CELLS test hook (256-cell passes over 200-300 cells, so the last pass is ragged).  Same bins up to the order
    of the additions."""
    case = cases.SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    with harness.open_session(str(tmp_path), case, surf) as h:
        one, _ = h.abi_spectra()
        monkeypatch.setenv("IS3D_PASS_CELLS", "256")
        many, st = h.abi_spectra()
    assert st.cells_total == len(surf["tau"])
    harness.assert_spectra_close(many, one, rtol=1e-12, what=name + " multi-pass vs single pass")
    harness.assert_spectra_close(many, ref, what=name + " multi-pass vs reference")
