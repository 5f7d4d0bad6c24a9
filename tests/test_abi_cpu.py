"""CPU checks: the C-ABI libraries load, export every declared symbol, and refuse to run without a GPU."""
import ctypes as C
import os
import re

import pytest

from is3d2_b200 import capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header: str):
    text = open(os.path.join(REPO, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(is3d_[a-z0-9_]+)\s*\(", text)))


def test_abi_exports_every_declared_symbol(libs):
    lib, host = libs
    for sym in _declared("is3d_b200.h"):
        assert hasattr(lib, sym), f"libis3d_b200.so does not export {sym}"
    assert sorted(capi.ABI_SYMBOLS) == _declared("is3d_b200.h")
    for sym in _declared("is3d_host.h"):
        assert hasattr(host, sym), f"libis3d_host.so does not export {sym}"
    assert sorted(capi.HOST_SYMBOLS) == _declared("is3d_host.h")


def test_params_struct_layout_matches_header():
    """ctypes mirror has the same field names, in order, as is3d_params / is3d_stats in the header."""
    text = open(os.path.join(REPO, "include", "is3d_b200.h")).read()
    for struct, cls in (("is3d_params", capi.Params), ("is3d_stats", capi.Stats), ("is3d_particle", capi.Particle)):
        body = re.search(r"typedef struct \{([^}]*)\} " + struct + ";", text).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(int64_t|int32_t|int|double)\s+", "", decl)
            names += [n.strip() for n in decl.split(",")]
        assert names == [f for f, _ in cls._fields_], struct


def test_create_fails_loudly_without_gpu(libs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib, _ = libs
    p = capi.Params()
    lib.is3d_default_params(C.byref(p))
    ctx = C.c_void_p()
    st = lib.is3d_create(C.byref(p), C.byref(ctx))
    assert st != 0 and not ctx.value
    assert b"no CPU path" in lib.is3d_last_error(None) or b"sm_" in lib.is3d_last_error(None)


def _smash_species():
    """(mass, sign, baryon) of the 444 chosen SMASH species through the host layer's PDG reader (no GPU involved)."""
    import numpy as np
    toks = open(os.path.join(REPO, "data", "PDG", "pdg_smash.dat")).read().split()
    pdg, k = {}, 0
    while k < len(toks):
        mcid, mass, b, nd = int(toks[k]), float(toks[k + 2]), int(toks[k + 5]), int(toks[k + 11])
        k += 12 + 8 * nd
        pdg[mcid] = (mass, b)
        if b > 0:
            pdg[-mcid] = (mass, -b)
    chosen = [int(l.split()[0]) for l in open(os.path.join(REPO, "data", "PDG", "chosen_particles_smash.dat")) if l.strip()]
    mass = np.array([pdg[c][0] for c in chosen])
    baryon = np.array([float(pdg[c][1]) for c in chosen])
    sign = np.where(baryon.astype(int) % 2 == 0, -1.0, 1.0)
    return mass, sign, baryon


@pytest.mark.parametrize("include_baryon,R", [(1, 4), (1, 3), (0, 4), (1, 7)])
def test_species_groups_have_one_baryon_number(libs, include_baryon, R):
    """Host logic behind the kernels' thread groups (is3d_species_groups): every species lands in exactly one class with
    its (mass, sign[, baryon number]); every class sits in exactly one slot; the valid slots of a group share ONE baryon
    number; padding only at the end of a baryon-number run."""
    import numpy as np
    lib, _ = libs
    mass, sign, baryon = _smash_species()
    ns = len(mass)
    assert ns == 444
    class_of = np.zeros(ns, dtype=np.int32)
    slots = np.full(4 * ns, -7, dtype=np.int32)
    nclass = C.c_int()
    lib.is3d_species_groups.restype = C.c_int
    n = lib.is3d_species_groups(ns, mass.ctypes.data_as(C.c_void_p), sign.ctypes.data_as(C.c_void_p), baryon.ctypes.data_as(C.c_void_p),
                                include_baryon, R, class_of.ctypes.data_as(C.c_void_p), slots.ctypes.data_as(C.c_void_p), len(slots),
                                C.byref(nclass))
    assert n > 0 and n % R == 0
    nc = nclass.value
    assert nc == (193 if include_baryon else len({(m, s) for m, s in zip(mass, sign)}))
    # classes: same key <=> same class
    key = {}
    for s in range(ns):
        k = (mass[s], sign[s], baryon[s] if include_baryon else 0.0)
        assert key.setdefault(k, class_of[s]) == class_of[s]
    assert len(key) == nc and sorted(set(class_of)) == list(range(nc))
    # slots: a permutation of the classes plus padding
    sl = slots[:n]
    assert sorted(c for c in sl if c >= 0) == list(range(nc))
    cls_b = {class_of[s]: (baryon[s] if include_baryon else 0.0) for s in range(ns)}
    for g in range(n // R):
        grp = sl[g * R:(g + 1) * R]
        assert grp[0] >= 0
        assert len({cls_b[c] for c in grp if c >= 0}) == 1
        pad = [i for i, c in enumerate(grp) if c < 0]
        assert pad == list(range(R - len(pad), R))
    if include_baryon and R == 4:
        assert n // R == 50        # 19 meson + 15 baryon + 15 antibaryon groups + the deuteron
    # error paths: capacity, baryon number out of range
    assert lib.is3d_species_groups(ns, mass.ctypes.data_as(C.c_void_p), sign.ctypes.data_as(C.c_void_p), baryon.ctypes.data_as(C.c_void_p),
                                   include_baryon, R, class_of.ctypes.data_as(C.c_void_p), slots.ctypes.data_as(C.c_void_p), 3, C.byref(nclass)) == -2
    bad = baryon.copy(); bad[5] = 3.0
    rc = lib.is3d_species_groups(ns, mass.ctypes.data_as(C.c_void_p), sign.ctypes.data_as(C.c_void_p), bad.ctypes.data_as(C.c_void_p),
                                 1, R, class_of.ctypes.data_as(C.c_void_p), slots.ctypes.data_as(C.c_void_p), len(slots), C.byref(nclass))
    assert rc == -3


def test_charge_conjugate_pairs_layout(libs):
    """Host logic behind the PAIR launches of the spectra kernels (is3d_species_pairs): every class is either a single or a
    member of exactly one pair; the members of a pair have the same mass and statistics and opposite baryon number (b > 0
    first); groups carry one baryon number (singles) resp. one |b| (pairs); SMASH: 58 pairs + 77 singles."""
    import numpy as np
    lib, _ = libs
    mass, sign, baryon = _smash_species()
    ns, R = len(mass), 4
    class_of = np.zeros(ns, dtype=np.int32)
    slots = np.full(4 * ns, -7, dtype=np.int32)
    nclass = C.c_int()
    lib.is3d_species_groups.restype = C.c_int
    lib.is3d_species_groups(ns, mass.ctypes.data_as(C.c_void_p), sign.ctypes.data_as(C.c_void_p), baryon.ctypes.data_as(C.c_void_p),
                            1, R, class_of.ctypes.data_as(C.c_void_p), slots.ctypes.data_as(C.c_void_p), len(slots), C.byref(nclass))
    singles = np.full(4 * ns, -7, dtype=np.int32)
    pairs = np.full(8 * ns, -7, dtype=np.int32)
    n1, n2 = C.c_int(), C.c_int()
    lib.is3d_species_pairs.restype = C.c_int
    nc = lib.is3d_species_pairs(ns, mass.ctypes.data_as(C.c_void_p), sign.ctypes.data_as(C.c_void_p), baryon.ctypes.data_as(C.c_void_p), R,
                                singles.ctypes.data_as(C.c_void_p), len(singles), pairs.ctypes.data_as(C.c_void_p), len(pairs),
                                C.byref(n1), C.byref(n2))
    assert nc == nclass.value == 193
    s1, p2 = singles[:n1.value], pairs[:n2.value].reshape(-1, 2)
    assert n1.value % R == 0 and len(p2) % R == 0
    prop = {class_of[s]: (mass[s], sign[s], baryon[s]) for s in range(ns)}
    used = [c for c in s1 if c >= 0] + [c for c in p2.ravel() if c >= 0]
    assert sorted(used) == list(range(nc))                         # every class exactly once
    real = p2[p2[:, 0] >= 0]
    assert len(real) == 58 and len([c for c in s1 if c >= 0]) == 77
    for a, b in real:
        assert b >= 0 and prop[a][0] == prop[b][0] and prop[a][1] == prop[b][1] and prop[a][2] > 0 and prop[b][2] == -prop[a][2]
    assert np.all((p2[:, 0] >= 0) == (p2[:, 1] >= 0))
    for g in range(len(p2) // R):
        grp = p2[g * R:(g + 1) * R]
        assert grp[0, 0] >= 0 and len({prop[c][2] for c in grp[:, 0] if c >= 0}) == 1
    for g in range(len(s1) // R):
        grp = s1[g * R:(g + 1) * R]
        assert grp[0] >= 0 and len({prop[c][2] for c in grp if c >= 0}) == 1
    # a list without antiparticles has no pairs
    keep = baryon >= 0
    nc2 = lib.is3d_species_pairs(int(keep.sum()), np.ascontiguousarray(mass[keep]).ctypes.data_as(C.c_void_p),
                                 np.ascontiguousarray(sign[keep]).ctypes.data_as(C.c_void_p),
                                 np.ascontiguousarray(baryon[keep]).ctypes.data_as(C.c_void_p), R, singles.ctypes.data_as(C.c_void_p), len(singles),
                                 pairs.ctypes.data_as(C.c_void_p), len(pairs), C.byref(n1), C.byref(n2))
    assert nc2 > 0 and n2.value == 0


def test_launch_order_sorts_columns_by_transverse_mass(libs):
    """Host logic behind the dropping of negligible items (is3d_launch_order): the thread columns (group, pT node) are a permutation
    sorted by sqrt(min mass of the group^2 + pT^2); every block of 128 consecutive columns therefore starts at a transverse mass that
    is >= the start of the block before it; bin_row maps every (class, pT) bin to the block that holds its column."""
    import numpy as np
    lib, _ = libs
    mass, sign, baryon = _smash_species()
    ns, R, T = len(mass), 4, 128
    class_of = np.zeros(ns, dtype=np.int32)
    slots = np.full(4 * ns, -7, dtype=np.int32)
    nclass = C.c_int()
    lib.is3d_species_groups.restype = C.c_int
    n = lib.is3d_species_groups(ns, mass.ctypes.data_as(C.c_void_p), sign.ctypes.data_as(C.c_void_p), baryon.ctypes.data_as(C.c_void_p),
                                0, R, class_of.ctypes.data_as(C.c_void_p), slots.ctypes.data_as(C.c_void_p), len(slots), C.byref(nclass))
    assert n > 0
    sl = np.ascontiguousarray(slots[:n])
    nc = nclass.value
    class_mass = np.zeros(nc)
    for s_ in range(ns):
        class_mass[class_of[s_]] = mass[s_]
    pT = np.loadtxt(os.path.join(REPO, "data", "tables", "momentum", "pT_table.dat"))[:, 0].copy()
    NpT, ngroups = len(pT), n // R
    order = np.full(ngroups * NpT, -1, dtype=np.int32)
    bin_row = np.full(nc * NpT, -5, dtype=np.int32)
    lib.is3d_launch_order.restype = C.c_int
    ncol = lib.is3d_launch_order(n, sl.ctypes.data_as(C.c_void_p), R, nc, class_mass.ctypes.data_as(C.c_void_p), NpT,
                                 pT.ctypes.data_as(C.c_void_p), T, order.ctypes.data_as(C.c_void_p), bin_row.ctypes.data_as(C.c_void_p))
    assert ncol == ngroups * NpT
    assert sorted(order) == list(range(ncol))                      # a permutation of the columns
    grp, ip = order // NpT, order % NpT
    m_min = np.array([min(class_mass[c] for c in sl[g * R:(g + 1) * R] if c >= 0) for g in range(ngroups)])
    key = np.sqrt(m_min[grp] ** 2 + pT[ip] ** 2)
    assert np.all(np.diff(key) >= 0.0)                             # sorted by the column's smallest mT
    starts = key[::T]
    assert np.all(np.diff(starts) >= 0.0) and starts[0] == pytest.approx(class_mass.min(), abs=pT.min() + 1e-12)
    # the widest block spans far less than the whole table: that is what lets a block agree on negligible cells
    widths = [key[b * T:(b + 1) * T].max() / key[b * T:(b + 1) * T].min() for b in range((ncol + T - 1) // T) if b > 0]
    assert max(widths) < 1.5
    # every bin of a class that sits in a slot belongs to the block of its column
    pos = {int(c): k for k, c in enumerate(order)}
    for g in range(ngroups):
        for c in sl[g * R:(g + 1) * R]:
            if c >= 0:
                for i in (0, NpT // 2, NpT - 1):
                    assert bin_row[c * NpT + i] == pos[g * NpT + i] // T
    assert lib.is3d_launch_order(n + 1, sl.ctypes.data_as(C.c_void_p), R, nc, class_mass.ctypes.data_as(C.c_void_p), NpT,
                                 pT.ctypes.data_as(C.c_void_p), T, order.ctypes.data_as(C.c_void_p), bin_row.ctypes.data_as(C.c_void_p)) == -1
