"""CPU checks of the C++ host layer against the reference (oracle/_ref, when built) and the data tables."""
import os
import subprocess

import numpy as np
import pytest

import refrun
from is3d_b200 import HostSession, synthetic, workdir


@pytest.mark.parametrize("hrg_eos,npdg", [(1, 327), (2, 493)])
def test_pdg_species_counts(libs, tmp_path, hrg_eos, npdg):
    # species counts of the reference's readers (UrQMD 327, SMASH 493 including antibaryons)
    root = workdir.make_workdir(str(tmp_path), dict(hrg_eos=hrg_eos, df_mode=2, dimension=3, mode=1), chosen="pikp")
    synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), synthetic.s3d(5, seed=1))
    with HostSession(root) as h:
        assert h.read_surface() == 5
        h.prepare_tables()
        pdg = h.pdg()
    assert pdg.shape == (npdg, 8)
    baryons = pdg[:, 3]
    assert (baryons > 0).sum() == (baryons < 0).sum()
    assert set(np.unique(pdg[:, 4])) == {-1.0, 1.0}


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("hrg_eos,df_mode,baryon", [(1, 1, 0), (2, 2, 0), (3, 4, 0), (2, 3, 0), (2, 2, 1), (2, 1, 1)])
def test_host_tables_match_reference(libs, tmp_path, hrg_eos, df_mode, baryon):
    """PDG parse, fast-mode densities, PTB tables, mode-1 reader and thermodynamic averages vs the reference."""
    s = synthetic.s3d(40, seed=21, baryon=bool(baryon))
    params = dict(operation=1, mode=1, hrg_eos=hrg_eos, dimension=3, df_mode=df_mode, include_baryon=baryon,
                  include_baryondiff_deltaf=baryon)
    chosen = "box" if hrg_eos == 3 else "pikp"
    ref = refrun.run_ref(str(tmp_path / "ref"), s, params, chosen=chosen, baryon=bool(baryon))
    root = workdir.make_workdir(str(tmp_path / "mine"), params, chosen=chosen)
    synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), s, baryon=bool(baryon))
    with HostSession(root) as h:
        h.read_surface()
        h.prepare_tables()
        pdg = h.pdg()
        ptb = h.ptb()
        seen = synthetic.roundtrip_mode1(s, baryon=bool(baryon))
        for k, name in enumerate(synthetic.SOA_COLUMNS):
            np.testing.assert_array_equal(h.surface_column(k), seen[name], err_msg=name)
    assert pdg.shape == ref["species"].shape
    np.testing.assert_array_equal(pdg[:, :5], ref["species"][:, :5])
    np.testing.assert_allclose(pdg[:, 5:], ref["species"][:, 5:], rtol=1e-13, atol=1e-300)
    if not baryon:
        j = ref["jonah"]
        np.testing.assert_allclose(ptb[0], j["bulkPi_over_P"], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(ptb[1], j["lambda2"], rtol=0, atol=0)
        np.testing.assert_allclose(ptb[2], j["z"], rtol=1e-13)
    a = open(tmp_path / "ref" / "tables" / "thermodynamic" / "average_thermodynamic_quantities.dat").read()
    b = open(tmp_path / "mine" / "tables" / "thermodynamic" / "average_thermodynamic_quantities.dat").read()
    assert a == b


def test_table_row_rule(libs, tmp_path):
    """A last line without a newline is not a row (reference Arsenal.cpp:79-125): surface.dat cell count."""
    root = workdir.make_workdir(str(tmp_path), dict(hrg_eos=1, df_mode=1, dimension=3, mode=1), chosen="pikp")
    p = os.path.join(root, "input", "surface.dat")
    synthetic.write_mode1(p, synthetic.s3d(4, seed=2))
    txt = open(p).read()
    open(p, "w").write(txt.rstrip("\n"))
    with HostSession(root) as h:
        assert h.read_surface() == 3


def test_missing_parameter_is_fatal(libs, tmp_path):
    """getVal on a missing key exits (reference ParameterReader.cpp:142-155); run in a subprocess."""
    root = workdir.make_workdir(str(tmp_path), dict(hrg_eos=1), chosen="pikp")
    p = os.path.join(root, "iS3D_parameters.dat")
    lines = [l for l in open(p) if not l.startswith("hrg_eos")]
    open(p, "w").writelines(lines)
    synthetic.write_mode1(os.path.join(root, "input", "surface.dat"), synthetic.s3d(2, seed=2))
    exe = os.path.join(workdir.REPO, "is3d_b200", "iS3D_b200.e")
    r = subprocess.run([exe], cwd=root, capture_output=True, text=True)
    assert r.returncode != 0
    assert "hrg_eos" in r.stdout and "not found" in r.stdout
