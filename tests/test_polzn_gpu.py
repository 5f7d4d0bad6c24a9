"""GPU parity of the spin-polarization path (K7) through the C ABI against the arrays of the unmodified
reference (tests/golden/make_golden_polzn.py) and the CPU oracle."""
import numpy as np
import pytest

import cases
import harness
import oracle_api
from is3d2_b200 import synthetic, workdir

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.POLZN_CASES))
def test_polarization_matches_reference(libs, tmp_path, monkeypatch, name):
    monkeypatch.setenv("IS3D_POLZN_CHUNK_COMPAT", "1")          # the reference's in-chunk vorticity index (opt-in)
    case = cases.POLZN_CASES[name]
    surf, vort, ref = harness.load_golden_polzn(name)
    with harness.open_session(str(tmp_path), case, surf) as h:
        h.abi_set_vorticity(vort)
        got, st = h.abi_polarization()
        again, _ = h.abi_polarization()
    got = np.stack(got)
    assert st.cells_total == len(surf["tau"]) and st.kernel_launches == 2
    np.testing.assert_array_equal(got, np.stack(again))                        # deterministic reduction
    worst = harness.assert_polzn_close(got, ref, what=name)
    # the physical observable: mean polarization vector S_mu / norm
    big = np.abs(ref[4]) > 1e-6 * np.abs(ref[4]).max()
    for k in range(4):
        np.testing.assert_allclose((got[k] / got[4])[big], (ref[k] / ref[4])[big], rtol=1e-9, atol=1e-14)
    print(f"{name}: max rel err {worst:.3e}")


def test_polarization_classes_and_corrected_index_match_oracle(libs, tmp_path, monkeypatch):
    """All 444 SMASH species (species classes by (mass, sign)) on a few cells against the CPU oracle, and the corrected
    vorticity index (polzn_chunk_compat = 0) on a surface with more than 10 000 cells."""
    case = dict(surface=("s3d", dict(n=9, seed=63)), params=cases._p(operation=1, mode=5, df_mode=2), chosen="smash")
    surf = synthetic.roundtrip_mode1(cases.make_surface(case["surface"]))
    vort = np.random.default_rng(5).uniform(-0.05, 0.05, (6, 9))
    with harness.open_session(str(tmp_path / "gpu"), case, surf) as h:
        h.abi_set_vorticity(vort)
        got, _ = h.abi_polarization()
    root = workdir.make_workdir(str(tmp_path / "oracle"), case["params"], chosen=case["chosen"])
    rc, want = oracle_api.OracleProblem(root, case["params"], surf).polarization(vort)
    assert rc == 0
    harness.assert_polzn_close(np.stack(got), want, what="smash 9 cells vs oracle")
    # corrected index
    name = "pol_s3d_10257cells"
    case = cases.POLZN_CASES[name]
    surf, vort, ref = harness.load_golden_polzn(name)
    monkeypatch.delenv("IS3D_POLZN_CHUNK_COMPAT", raising=False)  # library default: every cell reads its own vorticity
    with harness.open_session(str(tmp_path / "gpu2"), case, surf) as h:
        h.abi_set_vorticity(vort)
        got, _ = h.abi_polarization()
    root = workdir.make_workdir(str(tmp_path / "oracle2"), case["params"], chosen=case["chosen"])
    rc, want = oracle_api.OracleProblem(root, case["params"], surf).polarization(vort, chunk_compat=0)
    assert rc == 0
    harness.assert_polzn_close(np.stack(got), want, what="corrected vorticity index vs oracle")
    assert not np.allclose(np.stack(got)[0], ref[0], rtol=1e-6)


def test_polarization_needs_vorticity(libs, tmp_path):
    from is3d2_b200 import Is3dError
    name = "pol_s3d"
    surf, vort, _ = harness.load_golden_polzn(name)
    with harness.open_session(str(tmp_path), cases.POLZN_CASES[name], surf) as h:
        with pytest.raises(Is3dError, match="vorticity"):
            h.abi_polarization()
        with pytest.raises(Is3dError, match="same number of cells"):
            h.abi_set_vorticity(vort[:, :10])


@pytest.mark.parametrize("name", ["pol_s3d", "pol_s2d_phi48"])
def test_host_polarization_files_match_reference(libs, tmp_path, name):
    """Mode-5 surface.dat through the host layer (reader with vorticity columns, calculate_spectra, the polarization pass
    and write_polzn_vector_toFile): results/S{t,x,y,n}.dat equal the files the unmodified reference writes, including its
    storage-order / read-order index mismatch (polzn_file_compat, is3d2_b200/host/emission.cpp)."""
    import os
    case = cases.POLZN_CASES[name]
    surf = cases.make_surface(case["surface"])
    z = np.load(os.path.join(harness.GOLDEN, f"{name}.npz"))
    root = workdir.make_workdir(str(tmp_path), case["params"], chosen=case["chosen"], **case.get("tables", {}))
    synthetic.write_mode5(os.path.join(root, "input", "surface.dat"), surf, baryon=False, seed=len(surf["tau"]))
    from is3d2_b200 import HostSession
    with HostSession(root) as h:
        assert h.read_surface() == len(surf["tau"])
        h.prepare()
        h.run()
    for k, c in enumerate("txyn"):
        text = open(os.path.join(root, "results", f"S{c}.dat")).read()
        rows = np.array([[float(v) for v in l.split("\t")] for l in text.split("\n") if l])
        assert rows.shape == (z["files"].shape[1], 4)
        np.testing.assert_allclose(rows[:, 3], z["files"][k], rtol=2e-8, atol=1e-14, err_msg=f"S{c}.dat")
        assert text.count("\n\n") == rows.shape[0] // 51                      # blank line after every phi block of 51 pT rows
