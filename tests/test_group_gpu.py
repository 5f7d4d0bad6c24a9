"""Multi-GPU inside the product (include/is3d_b200.h "multi-GPU" section, csrc/comm.cu): cells sharded over the devices of
one process (is3d_group_*, driven here through the C++ host layer with IS3D_DEVICES) or over one context per rank
(is3d_comm_attach), combined by the product's own ncclAllReduce.  Sharded == unsharded is the property: spectra, dN/dX and
yields to 1e-12 (different summation order), sampled particle lists hadron for hadron.

The 2-device tests skip on a one-GPU box (run them with `gpurun --gpus 2`)."""
import ctypes as C
import threading

import numpy as np
import pytest

import cases
import harness
from is3d2_b200 import capi

pytestmark = pytest.mark.gpu


def _ndev() -> int:
    lib, _ = capi.load_libraries()
    lib.is3d_device_count.restype = C.c_int
    return int(lib.is3d_device_count())


need2 = pytest.mark.skipif("_ndev() < 2", reason="needs two GPUs (gpurun --gpus 2)")


def test_group_of_one_device_is_the_plain_context(libs, tmp_path, monkeypatch):
    """IS3D_DEVICES=0 goes through is3d_group_* with a single context: same bits as the golden-tested path."""
    name = "s3d_m2_baryon"
    case = cases.SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    monkeypatch.setenv("IS3D_DEVICES", "0")
    with harness.open_session(str(tmp_path), case, surf) as h:
        h.run()
        got = h.spectra()
        direct, _ = h.abi_spectra()
        assert h.lib.is3d_comm_size(h.ctx) == 1
    harness.assert_spectra_close(got, ref, what=name)
    np.testing.assert_array_equal(got, direct)


@need2
@pytest.mark.parametrize("name", ["s3d_m2_baryon", "s3d_m1_257cells", "s3d_m3", "s3d_m4", "s3d_m2_smash_17cells"])
def test_spectra_sharded_over_two_gpus(libs, tmp_path, monkeypatch, name):
    case = cases.SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    with harness.open_session(str(tmp_path / "one"), case, surf) as h:
        h.run()
        one = h.spectra()
    monkeypatch.setenv("IS3D_DEVICES", "0,1")
    with harness.open_session(str(tmp_path / "two"), case, surf) as h:
        h.run()
        two = h.spectra()
        st = h.stats()
        assert h.lib.is3d_comm_size(h.ctx) == 2 and h.lib.is3d_comm_collectives(h.ctx) >= 1
    assert st.cells_total == len(surf["tau"])
    harness.assert_spectra_close(two, one, rtol=1e-12, what=name + " 2 GPUs vs 1 GPU")
    harness.assert_spectra_close(two, ref, what=name + " 2 GPUs vs reference")


@need2
def test_famod_chain_free_sharded_over_two_gpus(libs, tmp_path, monkeypatch):
    """df_mode 5 in the library's default chain-free policy shards like every other mode (cells independent)."""
    name = "vah"
    case = cases.M5_CHAINFREE_CASES[name]
    surf, ref = harness.load_golden_m5free(name)
    monkeypatch.setenv("IS3D_DEVICES", "0-1")
    with harness.open_session(str(tmp_path), case, surf, famod_chain=0) as h:
        h.run()
        two = h.spectra()
    harness.assert_spectra_close(two, ref, what="chain-free df_mode 5 on 2 GPUs vs one-cell reference runs")


@need2
def test_dndx_and_yield_sharded_over_two_gpus(libs, tmp_path, monkeypatch):
    name = "dndx_s3d_m2_baryon"
    case = cases.DNDX_CASES[name]
    surf, _ = harness.load_golden_dndx(name)

    def run(root):
        with harness.open_session(root, case, surf) as h:
            h.run()
            tau, r, phi = C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
            ns = h.host.is3d_host_dndx(h.h, C.byref(tau), C.byref(r), C.byref(phi))
            import is3d2_b200.workdir as wd
            prm = wd.default_parameters()
            tb, rb, pb = int(float(prm["tau_bins"])), int(float(prm["r_bins"])), int(float(prm["phip_bins"]))
            return (np.ctypeslib.as_array(tau, shape=(ns, tb)).copy(), np.ctypeslib.as_array(r, shape=(ns, rb)).copy(),
                    np.ctypeslib.as_array(phi, shape=(ns, pb)).copy())
    one = run(str(tmp_path / "one"))
    monkeypatch.setenv("IS3D_DEVICES", "0,1")
    two = run(str(tmp_path / "two"))
    for a, b, k in zip(two, one, ("tau", "r", "phi")):
        harness.assert_hist_close(a, b, rtol=1e-12, what=f"dN/dX {k} 2 GPUs vs 1 GPU")
    # total yield through the sharded contexts
    sname = "smp_s3d_m3"
    scase = cases.SAMPLER_CASES[sname]
    ssurf, sref = harness.load_golden_sampler(sname)
    with harness.open_session(str(tmp_path / "y2"), scase, ssurf) as h:
        grp = h.host.is3d_host_group(h.h)
        assert h.lib.is3d_group_size(grp) == 2
        v, st = C.c_double(), capi.Stats()
        assert h.lib.is3d_group_total_yield(grp, C.byref(v), C.byref(st)) == 0, h.lib.is3d_group_last_error(grp)
    assert abs(v.value / float(sref["total_yield"]) - 1.0) < 1e-10
    assert st.cells_total == len(ssurf["tau"])


@need2
def test_sampler_sharded_over_two_gpus_gives_the_same_hadrons(libs, tmp_path, monkeypatch):
    """Philox streams keyed by the GLOBAL cell index: the merged event lists of a 2-GPU run are the single-GPU lists,
    hadron for hadron and in the same order."""
    name = "smp_s3d_m3"
    case = cases.SAMPLER_CASES[name]
    surf, ref = harness.load_golden_sampler(name)
    ov = dict(test_sampler=0, min_num_hadrons=40000.0, max_num_samples=2000.0)

    def run(root):
        with harness.open_session(root, case, surf, overrides=ov) as h:
            h.run()
            nev = h.host.is3d_host_events(h.h)
            out = []
            for e in range(nev):
                n = h.host.is3d_host_event_particles(h.h, e, None)
                a = np.zeros((n, 13))
                if n:
                    h.host.is3d_host_event_particles(h.h, e, a.ctypes.data_as(C.c_void_p))
                out.append(a)
            return out
    one = run(str(tmp_path / "one"))
    monkeypatch.setenv("IS3D_DEVICES", "0,1")
    two = run(str(tmp_path / "two"))
    assert len(one) == len(two) > 0 and sum(len(a) for a in one) > 10000
    for a, b in zip(one, two):
        np.testing.assert_array_equal(a, b)


@need2
def test_comm_attach_two_ranks_in_one_process(libs, tmp_path, monkeypatch):
    """The multi-process entry (is3d_comm_unique_id / is3d_comm_attach) exercised with two contexts on two devices driven by
    two host threads: each rank integrates its cell block, both receive the all-reduced spectra."""
    name = "s3d_m2_baryon"
    case = cases.SPECTRA_CASES[name]
    surf, ref = harness.load_golden(name)
    n = len(surf["tau"])
    cut = n // 2 + 7
    sessions = []
    for dev in (0, 1):
        monkeypatch.setenv("IS3D_DEVICE", str(dev))
        sessions.append(harness.open_session(str(tmp_path / f"r{dev}"), case, surf))
    lib = sessions[0].lib
    ident = (C.c_char * 128)()
    assert lib.is3d_comm_unique_id(ident) == 0
    blocks = [({k: v[:cut] for k, v in surf.items()}, 0), ({k: v[cut:] for k, v in surf.items()}, cut)]
    results, errors = [None, None], []

    def rank(i):
        try:
            h = sessions[i]
            h._check(lib.is3d_comm_attach(h.ctx, ident, 2, i), "is3d_comm_attach")
            h.abi_set_surface(blocks[i][0], global_offset=blocks[i][1])
            results[i] = h.abi_spectra()[0]
            lib.is3d_comm_detach(h.ctx)
        except Exception as e:  # noqa: BLE001
            errors.append(e)
    th = [threading.Thread(target=rank, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    for h in sessions:
        h.close()
    assert not errors, errors
    np.testing.assert_array_equal(results[0], results[1])            # every rank holds the same sum
    harness.assert_spectra_close(results[0], ref, what="2 ranks, product all-reduce, vs reference")


@need2
def test_sampler_self_test_histograms_sharded_over_two_gpus(libs, tmp_path, monkeypatch):
    """test_sampler = 1 through the host layer: the self-test histogram files of a 2-GPU run (per-device counters summed by
    is3d_group_sample_histograms) are the files of the 1-GPU run -- the same hadrons are sampled, counts are integers."""
    import os
    name = "smp_s3d_m3"
    case = cases.SAMPLER_CASES[name]
    surf, _ = harness.load_golden_sampler(name)
    ov = dict(test_sampler=1, min_num_hadrons=60000.0, max_num_samples=3000.0)

    def run(root):
        with harness.open_session(root, case, surf, overrides=ov) as h:
            h.run()
        out = {}
        base = os.path.join(root, "results", "sampled")
        for sub in sorted(os.listdir(base)):
            for f in sorted(os.listdir(os.path.join(base, sub))):
                out[sub + "/" + f] = open(os.path.join(base, sub, f)).read()
        return out
    one = run(str(tmp_path / "one"))
    monkeypatch.setenv("IS3D_DEVICES", "0,1")
    two = run(str(tmp_path / "two"))
    assert len(one) > 10 and sorted(one) == sorted(two)
    differ = [k for k in one if one[k] != two[k]]
    # the flow-coefficient files hold sums of cos / sin in FP64 (order of the atomic adds): compare those numerically
    for k in differ:
        assert k.startswith("vn/"), k
        a = np.array([[float(v) for v in l.split()] for l in one[k].splitlines() if l.strip() and l[0] != "#"])
        b = np.array([[float(v) for v in l.split()] for l in two[k].splitlines() if l.strip() and l[0] != "#"])
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
