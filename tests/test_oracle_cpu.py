"""The CPU oracle (oracle/cf_oracle.cpp) pinned against the golden vectors of the unmodified reference."""
import numpy as np
import pytest

import cases
import harness
import oracle_api
from is3d2_b200 import workdir

ORACLE_CASES = [n for n, c in cases.SPECTRA_CASES.items()
                if c["params"]["df_mode"] in (1, 2, 3, 4, 5) and c["chosen"] == "pikp" and c.get("tables") is None]


@pytest.mark.parametrize("name", ORACLE_CASES)
def test_oracle_matches_reference_golden(libs, tmp_path, name):
    case = cases.SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    root = workdir.make_workdir(str(tmp_path), case["params"], chosen=case["chosen"], **case.get("tables", {}))
    prob = oracle_api.OracleProblem(root, case["params"], surf)
    rc, got, st = prob.spectra()
    assert rc == 0
    # same loop order and expression association as the reference; what is left is compiler-level rounding
    # amplified in the few bins where positive and negative p.dsigma contributions cancel
    worst = harness.assert_spectra_close(got, ref, rtol=1e-11, what=name)
    print(name, worst)


@pytest.mark.parametrize("name", list(cases.DNDX_CASES))
def test_oracle_dndx_matches_reference_golden(libs, tmp_path, name):
    case = cases.DNDX_CASES[name]
    surf, ref = harness.load_golden_dndx(name)
    root = workdir.make_workdir(str(tmp_path), case["params"], chosen=case["chosen"], **case.get("tables", {}))
    prob = oracle_api.OracleProblem(root, case["params"], surf)
    rc, clean, st = prob.dndx()
    assert rc == 0
    got = harness.normalise_dndx({k: harness.emulate_partial_memset(v) for k, v in clean.items()}, case["params"])
    for k in ("tau", "r", "phi"):
        harness.assert_hist_close(got[k], ref[k], rtol=1e-11, what=f"{name}/{k}")


@pytest.mark.parametrize("name", list(cases.SAMPLER_CASES))
def test_oracle_total_yield_matches_reference(libs, tmp_path, name):
    """calculate_total_yield of the unmodified reference, dumped in full precision by oracle/ref_harness.cpp."""
    case = cases.SAMPLER_CASES[name]
    surf, ref = harness.load_golden_sampler(name)
    root = workdir.make_workdir(str(tmp_path), case["params"], chosen=case["chosen"])
    prob = oracle_api.OracleProblem(root, case["params"], surf)
    rc, ntot = prob.total_yield()
    assert rc == 0
    # with baryon diffusion the reference multiplies V.dsigma by an UNINITIALISED dsigma_space (ParticleSampler.cpp
    # :606-609 never calls compute_dsigma_magnitude); the diffusion term is ~1e-5 of the yield
    tol = 1e-4 if case["params"].get("include_baryondiff_deltaf") else 1e-12
    assert abs(ntot / float(ref["total_yield"]) - 1.0) < tol


@pytest.mark.parametrize("name", list(cases.POLZN_CASES))
def test_oracle_polarization_matches_reference_golden(libs, tmp_path, name):
    """cf_oracle_polarization against the arrays of the unmodified reference (Polarization.cpp), including the in-chunk
    vorticity index for the surface with more than 10 000 cells."""
    case = cases.POLZN_CASES[name]
    surf, vort, ref = harness.load_golden_polzn(name)
    root = workdir.make_workdir(str(tmp_path), case["params"], chosen=case["chosen"], **case.get("tables", {}))
    prob = oracle_api.OracleProblem(root, case["params"], surf)
    rc, got = prob.polarization(vort, chunk_compat=1)
    assert rc == 0
    worst = harness.assert_polzn_close(got, ref, rtol=1e-11, what=name)
    print(name, worst)
    if len(surf["tau"]) > 10000:
        rc, fixed = prob.polarization(vort, chunk_compat=0)
        assert rc == 0 and not np.allclose(fixed[0], ref[0], rtol=1e-6)      # the corrected index is a different result


@pytest.mark.parametrize("name", [n for n, c in cases.M5_CHAINFREE_CASES.items() if c["chosen"] == "pikp"])
def test_oracle_chain_free_matches_one_cell_reference_runs(libs, tmp_path, name):
    """df_mode 5 with every cell's Newton solve started from (T, 1, 1) (famod_chain = 0), against the sum of ONE-CELL runs
    of the unmodified reference -- a one-cell surface never has a previous solution (MomentumSpectra.cpp:1308-1313), so
    those runs are the chain-free policy of the reference itself."""
    case = cases.M5_CHAINFREE_CASES[name]
    surf, ref = harness.load_golden_m5free(name)
    root = workdir.make_workdir(str(tmp_path), case["params"], chosen=case["chosen"], **case.get("tables", {}))
    rc, got, st = oracle_api.OracleProblem(root, case["params"], surf, famod_chain=0).spectra()
    assert rc == 0
    worst = harness.assert_spectra_close(got, ref, rtol=1e-11, what=name + " chain-free oracle vs one-cell reference runs")
    print(name, worst)
