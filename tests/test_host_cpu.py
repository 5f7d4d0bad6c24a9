"""CPU checks of the C++ host layer against the reference (oracle/_ref, when built) and the data tables."""
import os
import subprocess

import numpy as np
import pytest

import refrun
from is3d2_b200 import HostSession, synthetic, workdir


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
    exe = os.path.join(workdir.REPO, "is3d2_b200", "iS3D_b200.e")
    r = subprocess.run([exe], cwd=root, capture_output=True, text=True)
    assert r.returncode != 0
    assert "hrg_eos" in r.stdout and "not found" in r.stdout


def test_parallel_reader_is_thread_count_independent(libs, tmp_path, monkeypatch):
    """The multi-threaded surface.dat parser delivers the same flat number stream as the serial
    `ifstream >> double` loop of the reference: identical columns for 1, 3 and 8 threads on a file whose slices cut
    through rows, with explicit '+' signs, exponents in both cases and irregular whitespace; every value equals
    float(token) (correct rounding)."""
    n = 60_000
    s = synthetic.s3d(n, seed=77, baryon=True)
    root = workdir.make_workdir(str(tmp_path), dict(hrg_eos=2, df_mode=2, dimension=3, mode=1, include_baryon=1), chosen="pikp")
    p = os.path.join(root, "input", "surface.dat")
    synthetic.write_mode1(p, s, baryon=True)
    lines = open(p).read().split("\n")
    # decorate a few rows: '+' sign, upper-case exponent, tabs and runs of blanks
    lines[5] = "\t".join("+" + t if not t.startswith("-") else t for t in lines[5].split())
    lines[6] = "   ".join(t.replace("e", "E") for t in lines[6].split()) + "   "
    open(p, "w").write("\n".join(lines))
    seen = synthetic.roundtrip_mode1(s, baryon=True)
    cols = {}
    for nt in (1, 3, 8):
        monkeypatch.setenv("IS3D_READER_THREADS", str(nt))
        with HostSession(root) as h:
            assert h.read_surface() == n
            cols[nt] = [h.surface_column(k).copy() for k in range(len(synthetic.SOA_COLUMNS))]
    for k, name in enumerate(synthetic.SOA_COLUMNS):
        np.testing.assert_array_equal(cols[1][k], seen[name], err_msg=name)
        np.testing.assert_array_equal(cols[3][k], cols[1][k], err_msg=name)
        np.testing.assert_array_equal(cols[8][k], cols[1][k], err_msg=name)


def test_reader_stops_at_a_malformed_token(libs, tmp_path, monkeypatch):
    """`ifstream >> double` fails at the first non-numeric token and every later value stays 0 (the reference keeps
    looping with the stream in its failed state); the parallel parser reproduces that for any thread count."""
    n = 4000
    s = synthetic.s3d(n, seed=78)
    root = workdir.make_workdir(str(tmp_path), dict(hrg_eos=2, df_mode=2, dimension=3, mode=1), chosen="pikp")
    p = os.path.join(root, "input", "surface.dat")
    synthetic.write_mode1(p, s)
    lines = open(p).read().split("\n")
    bad_row = 2500
    toks = lines[bad_row].split()
    toks[7] = "oops"
    lines[bad_row] = " ".join(toks)
    open(p, "w").write("\n".join(lines))
    seen = synthetic.roundtrip_mode1(s)
    for nt in (1, 5):
        monkeypatch.setenv("IS3D_READER_THREADS", str(nt))
        with HostSession(root) as h:
            assert h.read_surface() == n
            tau = h.surface_column(0).copy()
            dan = h.surface_column(7).copy()
            ux = h.surface_column(8).copy()
        np.testing.assert_array_equal(tau[:bad_row + 1], seen["tau"][:bad_row + 1])
        assert np.all(tau[bad_row + 1:] == 0.0)
        np.testing.assert_array_equal(dan[:bad_row], seen["dan"][:bad_row])
        assert np.all(dan[bad_row:] == 0.0) and np.all(ux[bad_row:] == 0.0)


def test_reader_uneven_rows_follow_the_stream(libs, tmp_path, monkeypatch):
    """Rows need not hold exactly `columns` numbers: the reference counts newline-terminated rows for the cell count and
    then reads one flat stream.  A row wrapped onto two lines gives one more (partly empty) cell; the parallel parser
    must notice that its per-slice offset guess is off and fall back to exact token offsets."""
    n = 30_000
    s = synthetic.s3d(n, seed=79)
    root = workdir.make_workdir(str(tmp_path), dict(hrg_eos=2, df_mode=2, dimension=3, mode=1), chosen="pikp")
    p = os.path.join(root, "input", "surface.dat")
    synthetic.write_mode1(p, s)
    lines = open(p).read().split("\n")
    toks = lines[100].split()
    lines[100] = " ".join(toks[:7]) + "\n" + " ".join(toks[7:])           # wrapped early in the file
    open(p, "w").write("\n".join(lines))
    flat = np.array(open(p).read().split(), dtype=np.float64)
    want_tau = flat[0::20]
    got = {}
    for nt in (1, 6):
        monkeypatch.setenv("IS3D_READER_THREADS", str(nt))
        with HostSession(root) as h:
            assert h.read_surface() == n + 1
            got[nt] = h.surface_column(0).copy()
    np.testing.assert_array_equal(got[1][:n], want_tau[:n])
    assert got[1][n] == 0.0                                               # the stream ran dry before the extra cell
    np.testing.assert_array_equal(got[6], got[1])


@pytest.mark.skipif(not refrun.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("mode,dimension,baryon", [(1, 3, 1), (5, 3, 0), (5, 3, 1), (6, 3, 0), (6, 3, 1), (6, 2, 0), (7, 2, 0)])
def test_surface_readers_match_reference(libs, tmp_path, mode, dimension, baryon):
    """Column contracts and unit conversions of the four surface.dat layouts (reference
    readindata.cpp:167-729): every cell field and the thermodynamic-average side file equal what the unmodified
    reference's reader produces from the same text, bit for bit."""
    n = 257
    s = synthetic.s3d(n, seed=90 + mode, baryon=bool(baryon), dimension=dimension)
    params = dict(operation=1, mode=mode, hrg_eos=2, dimension=dimension, df_mode=2, include_baryon=baryon,
                  include_baryondiff_deltaf=baryon)
    roots = {}
    for who in ("ref", "mine"):
        root = workdir.make_workdir(str(tmp_path / who), params, chosen="pikp")
        p = os.path.join(root, "input", "surface.dat")
        if mode == 1:
            synthetic.write_mode1(p, s, baryon=bool(baryon))
        elif mode == 5:
            synthetic.write_mode5(p, s, baryon=bool(baryon), seed=3)
        elif mode == 6:
            synthetic.write_mode6(p, s, baryon=bool(baryon))
        else:
            synthetic.write_mode7(p, s)
        roots[who] = root
    ref = refrun.ref_surface(roots["ref"])
    with HostSession(roots["mine"]) as h:
        assert h.read_surface() == n
        mine = np.stack([h.surface_column(k) for k in range(25)], axis=1)
    assert ref.shape == (n, 31)
    # without baryon columns the reference leaves muB, nB, V^mu of its `new FO_surf[]` uninitialised (modes 1, 5); the
    # MUSIC and HIC-EventGen layouts always carry muB
    ncmp = 25 if baryon else (21 if mode in (6, 7) else 20)
    for k, name in enumerate(synthetic.SOA_COLUMNS[:ncmp]):
        np.testing.assert_array_equal(mine[:, k], ref[:, k], err_msg=f"mode {mode}: {name}")
    a = open(os.path.join(roots["ref"], "tables", "thermodynamic", "average_thermodynamic_quantities.dat")).read()
    b = open(os.path.join(roots["mine"], "tables", "thermodynamic", "average_thermodynamic_quantities.dat")).read()
    assert a == b


def test_surface_binary_cache(libs, tmp_path, monkeypatch):
    """IS3D_SURFACE_CACHE=1: the first read parses the text and writes input/surface.dat.soa, the second read takes the
    columns from it (bit-identical surface and average file); touching surface.dat or changing a reader setting
    invalidates the cache."""
    n = 5000
    s = synthetic.s3d(n, seed=81, baryon=True)
    params = dict(hrg_eos=2, df_mode=2, dimension=3, mode=1, include_baryon=1)
    root = workdir.make_workdir(str(tmp_path), params, chosen="pikp")
    p = os.path.join(root, "input", "surface.dat")
    synthetic.write_mode1(p, s, baryon=True)
    avg_file = os.path.join(root, "tables", "thermodynamic", "average_thermodynamic_quantities.dat")

    def read():
        with HostSession(root) as h:
            assert h.read_surface() == n
            return [h.surface_column(k).copy() for k in range(25)], open(avg_file).read()
    plain, avg0 = read()
    assert not os.path.exists(p + ".soa")
    monkeypatch.setenv("IS3D_SURFACE_CACHE", "1")
    first, avg1 = read()
    assert os.path.exists(p + ".soa") and os.path.getsize(p + ".soa") > 25 * 8 * n
    stamp = os.path.getmtime(p + ".soa")
    second, avg2 = read()
    assert os.path.getmtime(p + ".soa") == stamp                           # served from the cache, not rewritten
    for a, b, c in zip(plain, first, second):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)
    assert avg0 == avg1 == avg2
    # a changed surface.dat (new cell values, same length) must not be served from the stale cache
    s2 = synthetic.s3d(n, seed=82, baryon=True)
    synthetic.write_mode1(p, s2, baryon=True)
    os.utime(p, (stamp + 10, stamp + 10))
    third, _ = read()
    np.testing.assert_array_equal(third[0], synthetic.roundtrip_mode1(s2, baryon=True)["tau"])


def test_two_live_sessions_keep_their_own_roots(libs, tmp_path):
    """Two host sessions alive at once, used in turn: every entry point resolves the reference's fixed relative paths
    against ITS session's root (the thermodynamic-average side file, the tables), not against the root opened last."""
    roots, sessions, surfs = [], [], []
    for k, (seed, T_shift) in enumerate([(5, 0.0), (6, 0.004)]):
        root = workdir.make_workdir(str(tmp_path / f"s{k}"), dict(hrg_eos=2, df_mode=3, dimension=3, mode=1), chosen="pikp")
        s = synthetic.s3d(50, seed=seed)
        s["T"] = s["T"] + T_shift
        roots.append(root); surfs.append(s)
    a = HostSession(roots[0])
    a.set_surface(surfs[0])
    b = HostSession(roots[1])            # opened while a is alive
    b.set_surface(surfs[1])
    a.set_thermo_averages(np.array([0.151, 0.25, 0.042, 0.0, 0.0]))       # must land in a's directory, not b's
    a.prepare_tables()
    b.prepare_tables()
    pa, pb = a.pdg(), b.pdg()
    a.close(); b.close()
    fa = open(os.path.join(roots[0], "tables", "thermodynamic", "average_thermodynamic_quantities.dat")).read().split()
    fb = open(os.path.join(roots[1], "tables", "thermodynamic", "average_thermodynamic_quantities.dat")).read().split()
    assert abs(float(fa[0]) - 0.151) < 1e-12
    assert abs(float(fb[0]) - 0.151) > 1e-4                                 # b kept the averages of its own surface
    # fast-mode densities were evaluated at each session's own averages
    assert not np.allclose(pa[:, 5], pb[:, 5])
    with HostSession(roots[0]) as c:                                        # a fresh session on a's root reproduces a's densities
        c.set_surface(surfs[0])
        c.set_thermo_averages(np.array([0.151, 0.25, 0.042, 0.0, 0.0]))
        c.prepare_tables()
        np.testing.assert_array_equal(c.pdg()[:, 5], pa[:, 5])
