"""Parity cases shared by the golden-vector generator and the tests: (name, surface spec, parameters, species,
tables).  Sizes are chosen so the reference finishes each in about a second."""
from __future__ import annotations

import os

from is3d2_b200 import synthetic

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

BASE = dict(operation=1, mode=1, hrg_eos=2, dimension=3, include_baryon=0, include_bulk_deltaf=1,
            include_shear_deltaf=1, include_baryondiff_deltaf=0, regulate_deltaf=0, outflow=0)


def _p(**kw):
    d = dict(BASE)
    d.update(kw)
    return d


# name -> dict(surface=(kind, kwargs), params, chosen, tables)
SPECTRA_CASES = {
    # BASELINE.json config 1: the bundled static cell, pikp, Grad, UrQMD, 2+1d
    "bundled_m1_2d": dict(surface=("bundled", {}), params=_p(df_mode=1, hrg_eos=1, dimension=2), chosen="pikp"),
    "s3d_m1": dict(surface=("s3d", dict(n=300, seed=12345)), params=_p(df_mode=1), chosen="pikp"),
    "s3d_m2": dict(surface=("s3d", dict(n=300, seed=12345)), params=_p(df_mode=2), chosen="pikp"),
    "s3d_m2_phi48_reg_outflow": dict(surface=("s3d", dict(n=120, seed=5)), params=_p(df_mode=2, hrg_eos=1, regulate_deltaf=1, outflow=1),
                                     chosen="pikp", tables=dict(phi_table="phi_table_48pt.dat")),
    # BASELINE.json config 2 in miniature: RTA Chapman-Enskog with bulk + shear + baryon diffusion
    "s3d_m2_baryon": dict(surface=("s3d", dict(n=300, seed=7, baryon=True)),
                          params=_p(df_mode=2, include_baryon=1, include_baryondiff_deltaf=1), chosen="pikp"),
    "s3d_m1_baryon": dict(surface=("s3d", dict(n=300, seed=7, baryon=True)),
                          params=_p(df_mode=1, include_baryon=1, include_baryondiff_deltaf=1), chosen="pikp"),
    "s3d_m2_baryon_nodiff": dict(surface=("s3d", dict(n=200, seed=8, baryon=True)),
                                 params=_p(df_mode=2, include_baryon=1, include_baryondiff_deltaf=0), chosen="pikp"),
    "s2d_m1_phi48": dict(surface=("s3d", dict(n=100, seed=3, dimension=2)), params=_p(df_mode=1, hrg_eos=1, dimension=2),
                         chosen="pikp", tables=dict(phi_table="phi_table_48pt.dat")),
    "s2d_m2": dict(surface=("s3d", dict(n=150, seed=4, dimension=2)), params=_p(df_mode=2, dimension=2), chosen="pikp"),
    "s3d_m1_noshear_nobulk": dict(surface=("s3d", dict(n=100, seed=9)), params=_p(df_mode=1, include_bulk_deltaf=0, include_shear_deltaf=0), chosen="pikp"),
    # ragged sizes around the 256-cell tile / chunk boundaries, all SMASH species on a few cells
    "s3d_m2_smash_17cells": dict(surface=("s3d", dict(n=17, seed=11)), params=_p(df_mode=2), chosen="smash"),
    "s3d_m1_257cells": dict(surface=("s3d", dict(n=257, seed=12)), params=_p(df_mode=1), chosen="pikp"),
    # ---- K2: PTM / PTB modified equilibrium (df_mode 3, 4); "stress" surfaces contain breakdown and pl < 0 cells ----
    "s3d_m3": dict(surface=("s3d", dict(n=300, seed=12345, stress=0.3)), params=_p(df_mode=3), chosen="pikp"),
    "s3d_m4": dict(surface=("s3d", dict(n=300, seed=12345, stress=0.3)), params=_p(df_mode=4), chosen="pikp"),
    "s3d_m3_calm": dict(surface=("s3d", dict(n=300, seed=21)), params=_p(df_mode=3), chosen="pikp"),
    "s3d_m4_calm": dict(surface=("s3d", dict(n=300, seed=21)), params=_p(df_mode=4), chosen="pikp"),
    "s3d_m3_phi48_reg_outflow": dict(surface=("s3d", dict(n=100, seed=5, stress=0.3)),
                                     params=_p(df_mode=3, hrg_eos=1, regulate_deltaf=1, outflow=1, deta_min=0.01),
                                     chosen="pikp", tables=dict(phi_table="phi_table_48pt.dat")),
    "s3d_m4_nobulk": dict(surface=("s3d", dict(n=200, seed=6, stress=0.3)), params=_p(df_mode=4, include_bulk_deltaf=0), chosen="pikp"),
    "s3d_m3_noshear": dict(surface=("s3d", dict(n=200, seed=6, stress=0.3)), params=_p(df_mode=3, include_shear_deltaf=0), chosen="pikp"),
    "s3d_m3_baryon": dict(surface=("s3d", dict(n=300, seed=7, baryon=True, stress=0.3)),
                          params=_p(df_mode=3, include_baryon=1, include_baryondiff_deltaf=1), chosen="pikp"),
    "s3d_m3_baryon_nodiff": dict(surface=("s3d", dict(n=200, seed=8, baryon=True, stress=0.3)),
                                 params=_p(df_mode=3, include_baryon=1, include_baryondiff_deltaf=0), chosen="pikp"),
    "s2d_m3": dict(surface=("s3d", dict(n=150, seed=3, dimension=2, stress=0.3)), params=_p(df_mode=3, dimension=2, hrg_eos=1), chosen="pikp"),
    "s2d_m4_box_phi48": dict(surface=("s3d", dict(n=40, seed=3, dimension=2, stress=0.3)), params=_p(df_mode=4, dimension=2, hrg_eos=3),
                             chosen=os.path.join(_GOLDEN, "chosen_box_small.dat"), tables=dict(phi_table="phi_table_48pt.dat")),
    "s3d_m3_smash_9cells": dict(surface=("s3d", dict(n=9, seed=13, stress=0.3)), params=_p(df_mode=3), chosen="smash"),
    # all SMASH species WITH baryon terms: mesons, baryons, antibaryons and the deuteron (b = 0, +1, -1, 2) -- the thread
    # groups of one baryon number and the exp(-b alpha_B) slots of the spectra kernels (spectra_df.cuh DfItemU)
    "s3d_m1_smash_baryon": dict(surface=("s3d", dict(n=11, seed=71, baryon=True)),
                                params=_p(df_mode=1, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
    "s3d_m2_smash_baryon": dict(surface=("s3d", dict(n=11, seed=72, baryon=True)),
                                params=_p(df_mode=2, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
    "s3d_m2_smash_baryon_reg_outflow": dict(surface=("s3d", dict(n=9, seed=73, baryon=True)),
                                            params=_p(df_mode=2, include_baryon=1, include_baryondiff_deltaf=1, regulate_deltaf=1, outflow=1),
                                            chosen="smash"),
    "s3d_m3_smash_baryon": dict(surface=("s3d", dict(n=11, seed=74, baryon=True, stress=0.3)),
                                params=_p(df_mode=3, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
    # ---- K3: PTMA modified anisotropic distribution (df_mode 5), reference-faithful initial-guess chain ----
    # BASELINE.json config 4: VAH-like surface (large P_L / P_T anisotropy) with df_mode 5
    "vah_m5": dict(surface=("s3d", dict(n=200, seed=51, vah=True)), params=_p(df_mode=5), chosen="pikp"),
    "s3d_m5": dict(surface=("s3d", dict(n=200, seed=52)), params=_p(df_mode=5), chosen="pikp"),
    "s3d_m5_stress_outflow": dict(surface=("s3d", dict(n=200, seed=53, stress=0.3)), params=_p(df_mode=5, outflow=1), chosen="pikp"),
    "s2d_m5_phi48": dict(surface=("s3d", dict(n=40, seed=54, dimension=2, vah=True)), params=_p(df_mode=5, dimension=2, hrg_eos=1),
                         chosen="pikp", tables=dict(phi_table="phi_table_48pt.dat")),
    "vah_m5_baryon": dict(surface=("s3d", dict(n=150, seed=55, vah=True, baryon=True)),
                          params=_p(df_mode=5, include_baryon=1, include_baryondiff_deltaf=1), chosen="pikp"),
    "vah_m5_smash_6cells": dict(surface=("s3d", dict(n=6, seed=56, vah=True)), params=_p(df_mode=5), chosen="smash"),
}


def make_surface(spec):
    kind, kw = spec
    if kind == "bundled":
        return synthetic.bundled_cell()
    if kind == "bench":                       # a cell range of the benchmark surface (synthetic.bench_surface)
        return synthetic.bench_surface(kw["begin"], kw["end"], baryon=kw.get("baryon", True))
    return synthetic.s3d(**kw)


# BASELINE.json config 2 / 5 at launch-realistic size: the first 2304 cells of the benchmark surface (9 tiles of 256 cells:
# several cell chunks x 10 column slices x 21 rapidity blocks, the launch shape bench.py times), all 444 SMASH species,
# df_mode 2 with bulk + shear + baryon diffusion.  One serial run of the unmodified reference (~90 s).
BIG_SPECTRA_CASES = {
    "bench_m2_smash_baryon_2304cells": dict(surface=("bench", dict(begin=0, end=2304)),
                                            params=_p(df_mode=2, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
    # the same for the modified-equilibrium kernels (K2): PTM with bulk, per-(cell, class) renormalisation, charge-conjugate pairs
    "bench_m3_smash_baryon_1024cells": dict(surface=("bench", dict(begin=0, end=1024)),
                                            params=_p(df_mode=3, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
}

# dN/dX at a size with several cell chunks per block column (GPU test only; ~1 min of serial reference)
BIG_DNDX_CASES = {
    "dndx_bench_m2_smash_baryon_512cells": dict(surface=("bench", dict(begin=0, end=512)),
                                                params=_p(operation=0, df_mode=2, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
}

# executable-level golden (tests/golden/make_golden_exe_tree.py): BASELINE.json config 5 in miniature -- MUSIC-format
# (mode 6) surface.dat of the benchmark surface's first 256 cells, all SMASH species, df_mode 2 + baryon diffusion
EXE_TREE_CASE = dict(surface=("bench", dict(begin=0, end=256)),
                     params=_p(mode=6, df_mode=2, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash")

# df_mode 5 under the chain-free initial-guess policy: goldens are sums of ONE-CELL runs of the unmodified reference
# (tests/golden/make_golden_m5_chainfree.py)
M5_CHAINFREE_CASES = {
    "vah": dict(surface=("s3d", dict(n=120, seed=151, vah=True)), params=_p(df_mode=5), chosen="pikp"),
    "s3d_stress_outflow": dict(surface=("s3d", dict(n=80, seed=153, stress=0.3)), params=_p(df_mode=5, outflow=1), chosen="pikp"),
    "vah_baryon": dict(surface=("s3d", dict(n=60, seed=155, vah=True, baryon=True)),
                       params=_p(df_mode=5, include_baryon=1, include_baryondiff_deltaf=1), chosen="pikp"),
    "s2d_phi48": dict(surface=("s3d", dict(n=24, seed=154, dimension=2, vah=True)), params=_p(df_mode=5, dimension=2, hrg_eos=1),
                      chosen="pikp", tables=dict(phi_table="phi_table_48pt.dat")),
    "vah_smash": dict(surface=("s3d", dict(n=8, seed=156, vah=True)), params=_p(df_mode=5), chosen="smash"),
}

# dN/dX (operation 0) parity cases: name -> same layout as SPECTRA_CASES (operation forced to 0)
DNDX_CASES = {
    "dndx_s3d_m1": dict(surface=("s3d", dict(n=200, seed=31)), params=_p(operation=0, df_mode=1), chosen="pikp"),
    "dndx_s3d_m2_baryon": dict(surface=("s3d", dict(n=200, seed=32, baryon=True)),
                               params=_p(operation=0, df_mode=2, include_baryon=1, include_baryondiff_deltaf=1), chosen="pikp"),
    "dndx_s2d_m2_phi48": dict(surface=("s3d", dict(n=60, seed=33, dimension=2)), params=_p(operation=0, df_mode=2, dimension=2, hrg_eos=1),
                              chosen="pikp", tables=dict(phi_table="phi_table_48pt.dat")),
    "dndx_s3d_m3": dict(surface=("s3d", dict(n=200, seed=34, stress=0.3)), params=_p(operation=0, df_mode=3), chosen="pikp"),
    "dndx_s3d_m4_reg_outflow": dict(surface=("s3d", dict(n=200, seed=35, stress=0.3)),
                                    params=_p(operation=0, df_mode=4, regulate_deltaf=1, outflow=1), chosen="pikp"),
    "dndx_s2d_m3": dict(surface=("s3d", dict(n=80, seed=36, dimension=2, stress=0.3)), params=_p(operation=0, df_mode=3, dimension=2), chosen="pikp"),
    # all SMASH species with baryon terms (b = 0, +1, -1, 2): uniform-baryon thread groups of the dN/dX kernels
    "dndx_s3d_m1_smash_baryon": dict(surface=("s3d", dict(n=12, seed=81, baryon=True)),
                                     params=_p(operation=0, df_mode=1, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
    "dndx_s3d_m2_smash_baryon": dict(surface=("s3d", dict(n=12, seed=82, baryon=True)),
                                     params=_p(operation=0, df_mode=2, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
    "dndx_s3d_m3_smash_baryon": dict(surface=("s3d", dict(n=12, seed=83, baryon=True, stress=0.3)),
                                     params=_p(operation=0, df_mode=3, include_baryon=1, include_baryondiff_deltaf=1), chosen="smash"),
    "dndx_s3d_m4_smash": dict(surface=("s3d", dict(n=12, seed=84, stress=0.3)), params=_p(operation=0, df_mode=4), chosen="smash"),
    "dndx_s2d_m4": dict(surface=("s3d", dict(n=80, seed=37, dimension=2, stress=0.3)), params=_p(operation=0, df_mode=4, dimension=2, hrg_eos=1), chosen="pikp"),
}

# sampler (operation 2) cases: the reference run behind each golden file samples ~2e6 hadrons with test_sampler = 1
_S = dict(operation=2, oversample=1, fast=1, test_sampler=1, sampler_seed=1, min_num_hadrons=2.0e6, max_num_samples=1.0e7)
SAMPLER_CASES = {
    "smp_s3d_m1": dict(surface=("s3d", dict(n=300, seed=41)), params=_p(df_mode=1, **_S), chosen="pikp"),
    "smp_s3d_m2_baryon": dict(surface=("s3d", dict(n=300, seed=42, baryon=True)),
                              params=_p(df_mode=2, include_baryon=1, include_baryondiff_deltaf=1, **_S), chosen="pikp"),
    "smp_s3d_m3": dict(surface=("s3d", dict(n=300, seed=43, stress=0.3)), params=_p(df_mode=3, **_S), chosen="pikp"),
    "smp_s3d_m4": dict(surface=("s3d", dict(n=300, seed=44, stress=0.3)), params=_p(df_mode=4, **_S), chosen="pikp"),
    "smp_s2d_m3": dict(surface=("s3d", dict(n=200, seed=45, dimension=2, stress=0.2)), params=_p(df_mode=3, dimension=2, hrg_eos=1, **_S), chosen="pikp"),
    # BASELINE.json config 3 exactly: full SMASH HRG, df_mode 3 (PTM), oversampled events
    "smp_s3d_m3_smash": dict(surface=("s3d", dict(n=200, seed=146, stress=0.3)), params=_p(df_mode=3, **_S), chosen="smash"),
    "smp_s3d_m2_smash": dict(surface=("s3d", dict(n=200, seed=46)), params=_p(df_mode=2, **_S), chosen="smash"),
    # fast = 0: species densities from the cell's own (T, muB) by 32-point Gauss-Laguerre sums (max_particle_number)
    "smp_s3d_m2_slow": dict(surface=("s3d", dict(n=300, seed=47)), params=_p(df_mode=2, **dict(_S, fast=0)), chosen="pikp"),
    "smp_s3d_m3_slow": dict(surface=("s3d", dict(n=300, seed=48, stress=0.3)), params=_p(df_mode=3, **dict(_S, fast=0)), chosen="pikp"),
    "smp_s3d_m4_slow_smash": dict(surface=("s3d", dict(n=200, seed=49, stress=0.3)), params=_p(df_mode=4, **dict(_S, fast=0)), chosen="smash"),
    "smp_s3d_m3_slow_baryon": dict(surface=("s3d", dict(n=300, seed=50, baryon=True, stress=0.3)),
                                   params=_p(df_mode=3, include_baryon=1, include_baryondiff_deltaf=1, **dict(_S, fast=0)), chosen="pikp"),
    # df_mode 5 (PTMA): sample_dN_pTdpTdphidy_famod
    "smp_vah_m5": dict(surface=("s3d", dict(n=300, seed=57, vah=True)), params=_p(df_mode=5, **_S), chosen="pikp"),
    "smp_s3d_m5_stress": dict(surface=("s3d", dict(n=300, seed=58, stress=0.3)), params=_p(df_mode=5, **_S), chosen="pikp"),
    "smp_s2d_m5": dict(surface=("s3d", dict(n=200, seed=59, dimension=2, vah=True)), params=_p(df_mode=5, dimension=2, hrg_eos=1, **_S), chosen="pikp"),
    # corners: 2+1d PTB with per-cell densities, Grad with the viscous corrections switched off, PTMA with a chemical potential
    "smp_s2d_m4_slow": dict(surface=("s3d", dict(n=200, seed=65, dimension=2, stress=0.2)), params=_p(df_mode=4, dimension=2, hrg_eos=1, **dict(_S, fast=0)), chosen="pikp"),
    "smp_s3d_m1_noshear_nobulk": dict(surface=("s3d", dict(n=300, seed=66)), params=_p(df_mode=1, include_bulk_deltaf=0, include_shear_deltaf=0, **_S), chosen="pikp"),
    "smp_vah_m5_baryon": dict(surface=("s3d", dict(n=300, seed=67, vah=True, baryon=True)),
                              params=_p(df_mode=5, include_baryon=1, include_baryondiff_deltaf=1, **_S), chosen="pikp"),
}

# spin polarization (mode-5 surfaces: thermal vorticity columns; runs after any operation, reference EmissionFunction.cpp:1304-1310)
_POL = dict(operation=1, mode=5, df_mode=2)
POLZN_CASES = {
    "pol_s3d": dict(surface=("s3d", dict(n=300, seed=61)), params=_p(**_POL), chosen="pikp"),
    "pol_s2d_phi48": dict(surface=("s3d", dict(n=120, seed=62, dimension=2)), params=_p(dimension=2, hrg_eos=1, **_POL), chosen="pikp",
                          tables=dict(phi_table="phi_table_48pt.dat")),
    # more than one 10 000-cell chunk: the reference pairs cell 10 000 + k with the vorticity of cell k (Polarization.cpp:125-130)
    "pol_s3d_10257cells": dict(surface=("s3d", dict(n=10257, seed=64)), params=_p(**_POL), chosen="pikp"),
}
