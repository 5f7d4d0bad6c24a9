"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (the CPU restatement, oracle/cf_oracle.cpp) and a
numpy loader that builds its inputs straight from a working directory's table files."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from is3d2_b200 import HostSession, workdir

ORACLE_DIR = os.path.join(workdir.REPO, "oracle")
_lib = None


class CfParams(C.Structure):
    _fields_ = [("operation", C.c_int), ("dimension", C.c_int), ("df_mode", C.c_int), ("include_baryon", C.c_int),
                ("include_bulk_deltaf", C.c_int), ("include_shear_deltaf", C.c_int), ("include_baryondiff_deltaf", C.c_int),
                ("regulate_deltaf", C.c_int), ("outflow", C.c_int), ("deta_min", C.c_double), ("mass_pion0", C.c_double),
                ("fast", C.c_int), ("y_cut", C.c_double), ("tau_min", C.c_double), ("tau_max", C.c_double),
                ("tau_bins", C.c_int), ("r_min", C.c_double), ("r_max", C.c_double), ("r_bins", C.c_int),
                ("phip_bins", C.c_int), ("famod_chain", C.c_int)]


dp = C.POINTER(C.c_double)


class CfInputs(C.Structure):
    _fields_ = [("n_cells", C.c_long), ("col", dp * 25), ("n_species", C.c_int), ("mass", dp), ("sign", dp),
                ("degeneracy", dp), ("baryon", dp), ("equilibrium_density", dp), ("bulk_density", dp),
                ("diffusion_density", dp), ("n_pdg", C.c_int), ("pdg_mass", dp), ("pdg_sign", dp),
                ("pdg_degeneracy", dp), ("pdg_baryon", dp), ("n_pT", C.c_int), ("n_phi", C.c_int), ("n_y", C.c_int),
                ("n_eta", C.c_int), ("pT", dp), ("pT_w", dp), ("phi", dp), ("phi_w", dp), ("y", dp), ("y_w", dp),
                ("eta", dp), ("eta_w", dp), ("n_alpha", C.c_int), ("n_gla", C.c_int), ("gla_root", dp),
                ("gla_weight", dp), ("n_T", C.c_int), ("n_muB", C.c_int), ("T_arr", dp), ("muB_arr", dp),
                ("c0", dp), ("c1", dp), ("c2", dp), ("c3", dp), ("c4", dp), ("F", dp), ("G", dp), ("betabulk", dp),
                ("betaV", dp), ("betapi", dp), ("n_ptb", C.c_int), ("ptb_x", dp), ("ptb_lambda2", dp), ("ptb_z", dp),
                ("ptb_x_max", C.c_double), ("T_avg", C.c_double), ("E_avg", C.c_double), ("P_avg", C.c_double),
                ("muB_avg", C.c_double), ("nB_avg", C.c_double)]


class CfStats(C.Structure):
    _fields_ = [("cells_skipped", C.c_long), ("cells_breakdown", C.c_long), ("cells_pl_negative", C.c_long),
                ("reconstruction_failures", C.c_long), ("newton_iterations", C.c_long), ("cells_out_of_table", C.c_long)]


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "cf_oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-f", os.path.join(ORACLE_DIR, "Makefile"), "oracle"])
        _lib = C.CDLL(so)
        _lib.cf_oracle_spectra.argtypes = [C.POINTER(CfParams), C.POINTER(CfInputs), C.c_void_p, C.POINTER(CfStats)]
        _lib.cf_oracle_dndx.argtypes = [C.POINTER(CfParams), C.POINTER(CfInputs), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(CfStats)]
        _lib.cf_oracle_total_yield.argtypes = [C.POINTER(CfParams), C.POINTER(CfInputs), dp]
        _lib.cf_oracle_cell_yields.argtypes = [C.POINTER(CfParams), C.POINTER(CfInputs), C.c_void_p, C.c_void_p]
        _lib.cf_oracle_polarization.argtypes = [C.POINTER(CfParams), C.POINTER(CfInputs), C.POINTER(C.c_void_p), C.c_int] + [C.c_void_p] * 5
    return _lib


def _table(path):
    return np.loadtxt(path, ndmin=2)


def _df_table(path, include_baryon):
    with open(path) as f:
        n_T, n_muB = int(f.readline()), int(f.readline())
        f.readline()
        a = np.loadtxt(f, ndmin=2)
    if not include_baryon:
        n_muB = 1
    a = a[: n_T * n_muB]
    return a[:n_T, 0].copy(), a[::n_T, 1].copy(), np.ascontiguousarray(a[:, 2])


class OracleProblem:
    """Everything cf_oracle needs, loaded with numpy from a working directory (tables) and the host layer (PDG)."""

    def __init__(self, root: str, params: dict, surface: dict, after_surface=None, famod_chain: int = 1):
        """after_surface(session): optional hook between set_surface and the table stage (sharded runs install the
        whole-surface thermodynamic averages there)."""
        full = workdir.default_parameters()
        full.update({k: str(v) for k, v in params.items()})
        g = lambda k: float(full[k])  # noqa: E731
        self.p = CfParams()
        for name, _ in CfParams._fields_:
            if name == "famod_chain":
                self.p.famod_chain = famod_chain
            elif name in ("deta_min", "mass_pion0", "y_cut", "tau_min", "tau_max", "r_min", "r_max"):
                setattr(self.p, name, g(name))
            else:
                setattr(self.p, name, int(g(name)))
        baryon = bool(self.p.include_baryon)
        self.keep = []
        inp = CfInputs()

        def put(field, arr):
            a = np.ascontiguousarray(arr, dtype=np.float64)
            self.keep.append(a)
            setattr(inp, field, a.ctypes.data_as(dp))
            return a

        n = len(surface["tau"])
        inp.n_cells = n
        from is3d2_b200.capi import SURFACE_COLUMNS
        for k, name in enumerate(SURFACE_COLUMNS):
            a = np.ascontiguousarray(surface.get(name, np.zeros(n)), dtype=np.float64)
            self.keep.append(a)
            inp.col[k] = a.ctypes.data_as(dp)
        with HostSession(root) as h:
            h.set_surface(surface)
            if after_surface is not None:
                after_surface(h)
            h.prepare_tables()
            pdg = h.pdg()
            ptb = h.ptb()
        chosen = np.loadtxt(os.path.join(root, "PDG", "chosen_particles.dat"), ndmin=1).astype(np.int64)
        idx = [int(np.nonzero(pdg[:, 0] == m)[0][0]) for m in chosen]
        sp = pdg[idx]
        inp.n_species = len(idx)
        put("mass", sp[:, 1]); put("degeneracy", sp[:, 2]); put("baryon", sp[:, 3]); put("sign", sp[:, 4])
        put("equilibrium_density", sp[:, 5]); put("bulk_density", sp[:, 6]); put("diffusion_density", sp[:, 7])
        inp.n_pdg = len(pdg)
        put("pdg_mass", pdg[:, 1]); put("pdg_degeneracy", pdg[:, 2]); put("pdg_baryon", pdg[:, 3]); put("pdg_sign", pdg[:, 4])
        t = os.path.join(root, "tables")
        pT, phi = _table(os.path.join(t, "momentum", "pT_table.dat")), _table(os.path.join(t, "momentum", "phi_table.dat"))
        y, eta = _table(os.path.join(t, "momentum", "y_table.dat")), _table(os.path.join(t, "spacetime_rapidity", "eta_table.dat"))
        inp.n_pT, inp.n_phi, inp.n_y, inp.n_eta = len(pT), len(phi), len(y), len(eta)
        put("pT", pT[:, 0]); put("pT_w", pT[:, 1]); put("phi", phi[:, 0]); put("phi_w", phi[:, 1])
        put("y", y[:, 0]); put("y_w", y[:, 1]); put("eta", eta[:, 0]); put("eta_w", eta[:, 1])
        with open(os.path.join(t, "gauss", "gla_roots_weights.txt")) as f:
            n_alpha, n_gla = (int(v) for v in f.readline().split())
            gl = np.loadtxt(f)
        inp.n_alpha, inp.n_gla = n_alpha, n_gla
        put("gla_root", gl[:, 1]); put("gla_weight", gl[:, 2])
        eos = {1: "urqmd", 2: "smash", 3: "smash_box"}[int(g("hrg_eos"))]
        d = os.path.join(root, "deltaf_coefficients", "vh", eos)
        for name in ("c0", "c1", "c2", "c3", "c4", "F", "G", "betabulk", "betaV", "betapi"):
            T_arr, muB_arr, vals = _df_table(os.path.join(d, name + ".dat"), baryon)
            put(name, vals)
        inp.n_T, inp.n_muB = len(T_arr), len(muB_arr)
        put("T_arr", T_arr); put("muB_arr", muB_arr)
        if ptb is not None:
            inp.n_ptb = len(ptb[0])
            put("ptb_x", ptb[0]); put("ptb_lambda2", ptb[1]); put("ptb_z", ptb[2])
            inp.ptb_x_max = ptb[3]
        avg = [float(v) for v in open(os.path.join(t, "thermodynamic", "average_thermodynamic_quantities.dat")).read().split()]
        inp.T_avg, inp.E_avg, inp.P_avg, inp.muB_avg, inp.nB_avg = avg
        self.inp = inp
        self.ny = inp.n_y if self.p.dimension == 3 else 1

    def spectra(self):
        out = np.zeros((self.inp.n_species, self.inp.n_pT, self.inp.n_phi, self.ny))
        st = CfStats()
        rc = lib().cf_oracle_spectra(C.byref(self.p), C.byref(self.inp), out.ctypes.data, C.byref(st))
        return rc, out, st


def _dndx(self):
    ns = self.inp.n_species
    tau = np.zeros((ns, self.p.tau_bins)); r = np.zeros((ns, self.p.r_bins)); phi = np.zeros((ns, self.p.phip_bins))
    st = CfStats()
    rc = lib().cf_oracle_dndx(C.byref(self.p), C.byref(self.inp), tau.ctypes.data, r.ctypes.data, phi.ctypes.data, C.byref(st))
    return rc, {"tau": tau, "r": r, "phi": phi}, st


OracleProblem.dndx = _dndx


def _total_yield(self):
    v = C.c_double()
    rc = lib().cf_oracle_total_yield(C.byref(self.p), C.byref(self.inp), C.byref(v))
    return rc, v.value


def _cell_yields(self):
    n, ns = self.inp.n_cells, self.inp.n_species
    tot = np.zeros(n); lst = np.zeros((n, ns))
    rc = lib().cf_oracle_cell_yields(C.byref(self.p), C.byref(self.inp), tot.ctypes.data, lst.ctypes.data)
    return rc, tot, lst


OracleProblem.total_yield = _total_yield
OracleProblem.cell_yields = _cell_yields


def _polarization(self, vorticity, chunk_compat: int = 1):
    """(5, Ns, NpT, Nphi, Ny): St, Sx, Sy, Sn, Snorm in the spectra layout."""
    shape = (self.inp.n_species, self.inp.n_pT, self.inp.n_phi, self.ny)
    outs = [np.zeros(shape) for _ in range(5)]
    keep = [np.ascontiguousarray(w, dtype=np.float64) for w in vorticity]
    arr = (C.c_void_p * 6)(*[w.ctypes.data for w in keep])
    rc = lib().cf_oracle_polarization(C.byref(self.p), C.byref(self.inp), arr, chunk_compat, *[o.ctypes.data for o in outs])
    return rc, np.stack(outs)


OracleProblem.polarization = _polarization
