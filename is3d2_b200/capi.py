"""ctypes binding of include/is3d_b200.h (CUDA C ABI) and include/is3d_host.h (C++ host layer)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SURFACE_COLUMNS = ("tau", "x", "y", "eta", "dat", "dax", "day", "dan", "ux", "uy", "un", "E", "T", "P",
                   "pixx", "pixy", "pixn", "piyy", "piyn", "bulkPi", "muB", "nB", "Vx", "Vy", "Vn")


class Is3dError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("operation", C.c_int), ("dimension", C.c_int), ("df_mode", C.c_int), ("include_baryon", C.c_int),
                ("include_bulk_deltaf", C.c_int), ("include_shear_deltaf", C.c_int),
                ("include_baryondiff_deltaf", C.c_int), ("regulate_deltaf", C.c_int), ("outflow", C.c_int),
                ("deta_min", C.c_double), ("mass_pion0", C.c_double), ("fast", C.c_int), ("y_cut", C.c_double),
                ("sampler_seed", C.c_int64), ("test_sampler", C.c_int),
                ("pT_min", C.c_double), ("pT_max", C.c_double), ("pT_bins", C.c_int), ("y_bins", C.c_int),
                ("phip_bins", C.c_int), ("eta_cut", C.c_double), ("eta_bins", C.c_int),
                ("tau_min", C.c_double), ("tau_max", C.c_double), ("tau_bins", C.c_int),
                ("r_min", C.c_double), ("r_max", C.c_double), ("r_bins", C.c_int),
                ("device", C.c_int), ("famod_chain", C.c_int), ("dndx_bug_compat", C.c_int), ("polzn_chunk_compat", C.c_int),
                ("negligible_margin", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("cells_total", C.c_int64), ("cells_skipped", C.c_int64), ("cells_breakdown", C.c_int64),
                ("cells_pl_negative", C.c_int64), ("reconstruction_failures", C.c_int64),
                ("newton_iterations", C.c_int64), ("cells_out_of_table", C.c_int64),
                ("sampler_proposals", C.c_int64), ("sampler_accepted", C.c_int64),
                ("tau_breakdown", C.c_double), ("tau_pl_negative", C.c_double), ("kernel_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("evals_executed", C.c_int64), ("pair_evals_executed", C.c_int64),
                ("evals_dropped", C.c_int64), ("prune_reruns", C.c_int64)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


class Particle(C.Structure):
    _fields_ = [("chosen_index", C.c_int32), ("mcid", C.c_int32), ("event", C.c_int32), ("pad_", C.c_int32),
                ("mass", C.c_double), ("tau", C.c_double), ("x", C.c_double), ("y", C.c_double), ("eta", C.c_double),
                ("t", C.c_double), ("z", C.c_double), ("E", C.c_double), ("px", C.c_double), ("py", C.c_double),
                ("pz", C.c_double)]


PARTICLE_DTYPE = np.dtype([("chosen_index", "<i4"), ("mcid", "<i4"), ("event", "<i4"), ("pad_", "<i4"),
                           ("mass", "<f8"), ("tau", "<f8"), ("x", "<f8"), ("y", "<f8"), ("eta", "<f8"), ("t", "<f8"),
                           ("z", "<f8"), ("E", "<f8"), ("px", "<f8"), ("py", "<f8"), ("pz", "<f8")])

COMPACT_DTYPE = np.dtype([("chosen_index", "<i4"), ("event", "<i4"), ("tau", "<f8"), ("x", "<f8"), ("y", "<f8"), ("eta", "<f8"),
                          ("px", "<f8"), ("py", "<f8"), ("pz", "<f8")])

_lib = None
_host = None

# every symbol include/is3d_b200.h declares (the CPU test-suite checks the library exports all of them)
ABI_SYMBOLS = ["is3d_default_params", "is3d_create", "is3d_destroy", "is3d_last_error", "is3d_version", "is3d_device_count",
               "is3d_set_species", "is3d_set_pdg", "is3d_set_momentum_tables", "is3d_set_gauss_tables",
               "is3d_set_thermo_averages", "is3d_set_df_tables", "is3d_set_ptb_tables", "is3d_set_surface",
               "is3d_set_surface_device", "is3d_spectra_size", "is3d_spectra", "is3d_spectra_device", "is3d_dndx",
               "is3d_dndx_device", "is3d_total_yield", "is3d_cell_yields", "is3d_sample", "is3d_free_particles", "is3d_sample_compact", "is3d_sample_device", "is3d_expand_particles",
               "is3d_sample_histograms", "is3d_set_vorticity", "is3d_polarization", "is3d_measure_fp64_peak", "is3d_probe_math", "is3d_probe_aniso_math",
               "is3d_species_groups", "is3d_species_pairs", "is3d_launch_order", "is3d_stream", "is3d_copy_from_device",
               "is3d_comm_unique_id", "is3d_comm_attach", "is3d_comm_detach", "is3d_comm_size", "is3d_comm_collectives",
               "is3d_comm_last_error", "is3d_group_create", "is3d_group_destroy", "is3d_group_size", "is3d_group_ctx",
               "is3d_group_last_error", "is3d_group_cell_block", "is3d_group_set_surface", "is3d_group_set_vorticity",
               "is3d_group_spectra", "is3d_group_dndx", "is3d_group_total_yield", "is3d_group_polarization",
               "is3d_group_sample", "is3d_group_sample_compact", "is3d_group_sample_histograms"]
HOST_SYMBOLS = ["is3d_host_open", "is3d_host_close", "is3d_host_read_surface", "is3d_host_set_surface",
                "is3d_host_prepare", "is3d_host_prepare_tables", "is3d_host_context", "is3d_host_group", "is3d_host_run",
                "is3d_host_spectra", "is3d_host_dndx", "is3d_host_events", "is3d_host_event_particles",
                "is3d_host_seconds", "is3d_host_stats", "is3d_host_pdg", "is3d_host_ptb",
                "is3d_host_surface_column", "is3d_host_chosen", "is3d_host_thermo_sums", "is3d_host_set_thermo_averages"]


def load_libraries(libdir: str | None = None):
    """Load the in-tree shared libraries; fails loudly if they were not built (python -m is3d2_b200.build).
    libdir: directory holding another build of the two libraries (launch-shape variants under tools/; first call only)."""
    global _lib, _host
    if _lib is not None:
        return _lib, _host
    p = os.path.join(libdir or HERE, "libis3d_b200.so")
    ph = os.path.join(libdir or HERE, "libis3d_host.so")
    if not os.path.exists(p) or not os.path.exists(ph):
        raise Is3dError(f"{p} / {ph} not built: run `python -m is3d2_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    host = C.CDLL(ph, mode=C.RTLD_GLOBAL)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.is3d_last_error.restype = C.c_char_p
    lib.is3d_last_error.argtypes = [vp]
    lib.is3d_version.restype = C.c_char_p
    lib.is3d_default_params.argtypes = [C.POINTER(Params)]
    lib.is3d_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.is3d_destroy.argtypes = [vp]
    lib.is3d_set_surface.argtypes = [vp, C.c_int64, C.POINTER(vp), C.c_int64]
    lib.is3d_set_surface_device.argtypes = [vp, C.c_int64, C.POINTER(vp), C.c_int64]
    lib.is3d_spectra_size.restype = C.c_int64
    lib.is3d_spectra_size.argtypes = [vp]
    lib.is3d_spectra.argtypes = [vp, vp, C.POINTER(Stats)]
    lib.is3d_spectra_device.argtypes = [vp, vp, C.POINTER(Stats)]
    lib.is3d_dndx.argtypes = [vp, vp, vp, vp, C.POINTER(Stats)]
    lib.is3d_dndx_device.argtypes = [vp, vp, vp, vp, C.POINTER(Stats)]
    lib.is3d_total_yield.argtypes = [vp, dp, C.POINTER(Stats)]
    lib.is3d_cell_yields.argtypes = [vp, vp, vp, C.POINTER(Stats)]
    lib.is3d_sample.argtypes = [vp, C.c_int64, C.POINTER(vp), C.POINTER(C.c_int64), vp, C.POINTER(Stats)]
    lib.is3d_free_particles.argtypes = [vp]
    lib.is3d_free_particles.restype = None
    lib.is3d_sample_compact.argtypes = [vp, C.c_int64, C.POINTER(vp), C.POINTER(C.c_int64), vp, C.POINTER(Stats)]
    lib.is3d_sample_device.argtypes = [vp, C.c_int64, C.POINTER(vp), C.POINTER(C.c_int64), vp, C.POINTER(Stats)]
    lib.is3d_expand_particles.argtypes = [vp, vp, C.c_int64, vp]
    lib.is3d_sample_histograms.argtypes = [vp] + [vp] * 10
    lib.is3d_measure_fp64_peak.argtypes = [vp, dp]
    lib.is3d_set_vorticity.argtypes = [vp, C.c_int64, C.POINTER(vp)]
    lib.is3d_polarization.argtypes = [vp, vp, vp, vp, vp, vp, C.POINTER(Stats)]
    lib.is3d_probe_math.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
    lib.is3d_probe_aniso_math.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
    lib.is3d_copy_from_device.argtypes = [vp, vp, vp, C.c_size_t]
    lib.is3d_stream.restype = vp
    lib.is3d_stream.argtypes = [vp]
    host.is3d_host_open.restype = vp
    host.is3d_host_open.argtypes = [C.c_char_p, C.POINTER(C.c_char_p)]
    host.is3d_host_close.argtypes = [vp]
    host.is3d_host_read_surface.restype = C.c_int64
    host.is3d_host_read_surface.argtypes = [vp]
    host.is3d_host_set_surface.restype = C.c_int64
    host.is3d_host_set_surface.argtypes = [vp, C.c_int64, C.POINTER(vp)]
    host.is3d_host_prepare.argtypes = [vp]
    host.is3d_host_prepare_tables.argtypes = [vp]
    host.is3d_host_context.restype = vp
    host.is3d_host_context.argtypes = [vp]
    host.is3d_host_group.restype = vp
    host.is3d_host_group.argtypes = [vp]
    lib.is3d_device_count.restype = C.c_int
    lib.is3d_comm_unique_id.argtypes = [vp]
    lib.is3d_comm_attach.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.is3d_comm_detach.argtypes = [vp]
    lib.is3d_comm_detach.restype = None
    lib.is3d_comm_size.argtypes = [vp]
    lib.is3d_comm_collectives.restype = C.c_int64
    lib.is3d_comm_collectives.argtypes = [vp]
    lib.is3d_comm_last_error.restype = C.c_char_p
    lib.is3d_group_last_error.restype = C.c_char_p
    lib.is3d_group_last_error.argtypes = [vp]
    lib.is3d_group_size.argtypes = [vp]
    lib.is3d_group_total_yield.argtypes = [vp, dp, C.POINTER(Stats)]
    lib.is3d_group_spectra.argtypes = [vp, vp, C.POINTER(Stats)]
    host.is3d_host_run.argtypes = [vp]
    host.is3d_host_spectra.restype = C.c_int64
    host.is3d_host_spectra.argtypes = [vp, C.POINTER(dp), C.POINTER(C.c_int64)]
    host.is3d_host_dndx.restype = C.c_int64
    host.is3d_host_dndx.argtypes = [vp, C.POINTER(dp), C.POINTER(dp), C.POINTER(dp)]
    host.is3d_host_events.restype = C.c_int64
    host.is3d_host_events.argtypes = [vp]
    host.is3d_host_event_particles.restype = C.c_int64
    host.is3d_host_event_particles.argtypes = [vp, C.c_int64, vp]
    host.is3d_host_seconds.restype = C.c_double
    host.is3d_host_seconds.argtypes = [vp]
    host.is3d_host_stats.argtypes = [vp, C.POINTER(Stats)]
    host.is3d_host_pdg.restype = C.c_int64
    host.is3d_host_pdg.argtypes = [vp, vp]
    host.is3d_host_ptb.restype = C.c_int64
    host.is3d_host_ptb.argtypes = [vp, vp, vp, vp, dp]
    host.is3d_host_thermo_sums.argtypes = [vp, dp]
    host.is3d_host_set_thermo_averages.argtypes = [vp, dp]
    host.is3d_host_surface_column.restype = C.c_int64
    host.is3d_host_surface_column.argtypes = [vp, C.c_int, C.POINTER(dp)]
    host.is3d_host_chosen.restype = C.c_int64
    host.is3d_host_chosen.argtypes = [vp, ip]
    _lib, _host = lib, host
    return lib, host


def lib():
    return load_libraries()[0]


def host_lib():
    return load_libraries()[1]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _cols(surface: dict, as_device_ptrs: bool = False):
    arr = (C.c_void_p * 25)()
    keep = []
    for k, name in enumerate(SURFACE_COLUMNS):
        v = surface.get(name)
        if v is None:
            arr[k] = None
        elif as_device_ptrs:
            arr[k] = int(v)
        else:
            a = np.ascontiguousarray(v, dtype=np.float64)
            keep.append(a)
            arr[k] = a.ctypes.data
    return arr, keep


class HostSession:
    """One run of the host layer over a working directory laid out like the reference's repository root."""

    def __init__(self, root: str, overrides: dict | None = None):
        self.lib, self.host = load_libraries()
        ov = None
        if overrides:
            items = [f"{k} = {v}".encode() for k, v in overrides.items()]
            ov = (C.c_char_p * (len(items) + 1))(*items, None)
        self.h = self.host.is3d_host_open(root.encode(), ov)
        self.root = root
        self._keep = []

    def close(self):
        if self.h:
            self.host.is3d_host_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def read_surface(self) -> int:
        return self.host.is3d_host_read_surface(self.h)

    def set_surface(self, surface: dict) -> int:
        arr, keep = _cols(surface)
        n = len(keep[0])
        return self.host.is3d_host_set_surface(self.h, n, arr)

    def thermo_sums(self) -> np.ndarray:
        out = np.zeros(6)
        self.host.is3d_host_thermo_sums(self.h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def set_thermo_averages(self, avg5):
        a = np.ascontiguousarray(avg5, dtype=np.float64)
        assert a.size == 5
        self.host.is3d_host_set_thermo_averages(self.h, a.ctypes.data_as(C.POINTER(C.c_double)))

    def prepare_tables(self):
        self.host.is3d_host_prepare_tables(self.h)

    def prepare(self):
        self.host.is3d_host_prepare(self.h)

    @property
    def ctx(self):
        c = self.host.is3d_host_context(self.h)
        if not c:
            raise Is3dError("no CUDA context (call prepare())")
        return c

    def run(self):
        self.host.is3d_host_run(self.h)

    def _check(self, st: int, what: str):
        if st != 0:
            raise Is3dError(f"{what}: status {st}: {self.lib.is3d_last_error(self.ctx).decode()}")

    # ---- direct C-ABI calls on this session's context -----------------------------------------------------
    def abi_set_surface(self, surface: dict, global_offset: int = 0):
        arr, keep = _cols(surface)
        n = len(keep[0])
        self._check(self.lib.is3d_set_surface(self.ctx, n, arr, global_offset), "is3d_set_surface")

    def abi_set_surface_device(self, dev_ptrs: dict, n: int, global_offset: int = 0):
        arr, _ = _cols(dev_ptrs, as_device_ptrs=True)
        self._check(self.lib.is3d_set_surface_device(self.ctx, n, arr, global_offset), "is3d_set_surface_device")

    def spectra_shape(self):
        dims = (C.c_int64 * 4)()
        data = C.POINTER(C.c_double)()
        self.host.is3d_host_spectra(self.h, C.byref(data), dims)
        return tuple(int(d) for d in dims)

    def abi_spectra(self):
        n = self.lib.is3d_spectra_size(self.ctx)
        out = np.empty(n, dtype=np.float64)
        st = Stats()
        self._check(self.lib.is3d_spectra(self.ctx, _ptr(out), C.byref(st)), "is3d_spectra")
        return out.reshape(self.spectra_shape()), st

    def abi_spectra_device(self, dev_ptr: int):
        st = Stats()
        self._check(self.lib.is3d_spectra_device(self.ctx, dev_ptr, C.byref(st)), "is3d_spectra_device")
        return st

    def abi_fp64_peak(self) -> float:
        v = C.c_double()
        self._check(self.lib.is3d_measure_fp64_peak(self.ctx, C.byref(v)), "is3d_measure_fp64_peak")
        return v.value

    def abi_set_vorticity(self, w6):
        """w6: six arrays wtx wty wtn wxy wxn wyn (length = cells of the surface set last)."""
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in w6]
        arr = (C.c_void_p * 6)(*[a.ctypes.data for a in keep])
        self._check(self.lib.is3d_set_vorticity(self.ctx, len(keep[0]), arr), "is3d_set_vorticity")

    def abi_polarization(self):
        """(St, Sx, Sy, Sn, Snorm) in the spectra layout (Ns, NpT, Nphi, Ny), stats."""
        shape = self.spectra_shape()
        outs = [np.zeros(int(np.prod(shape))) for _ in range(5)]
        st = Stats()
        self._check(self.lib.is3d_polarization(self.ctx, *[_ptr(o) for o in outs], C.byref(st)), "is3d_polarization")
        return [o.reshape(shape) for o in outs], st

    def abi_probe_math(self, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.float64)
        e, r, s = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        self._check(self.lib.is3d_probe_math(self.ctx, x.size, _ptr(x), _ptr(e), _ptr(r), _ptr(s)), "is3d_probe_math")
        return e, r, s

    def abi_probe_aniso_math(self, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.float64)
        a, b, c = np.empty_like(x), np.empty_like(x), np.empty_like(x)
        self._check(self.lib.is3d_probe_aniso_math(self.ctx, x.size, _ptr(x), _ptr(a), _ptr(b), _ptr(c)), "is3d_probe_aniso_math")
        return a, b, c

    # ---- results kept by the host layer ------------------------------------------------------------------
    def spectra(self) -> np.ndarray:
        dims = (C.c_int64 * 4)()
        data = C.POINTER(C.c_double)()
        n = self.host.is3d_host_spectra(self.h, C.byref(data), dims)
        return np.ctypeslib.as_array(data, shape=(n,)).copy().reshape(tuple(int(d) for d in dims))

    def pdg(self) -> np.ndarray:
        n = self.host.is3d_host_pdg(self.h, None)
        out = np.empty((n, 8))
        self.host.is3d_host_pdg(self.h, _ptr(out))
        return out

    def chosen(self) -> np.ndarray:
        """MC ids of the chosen particles in output order."""
        n = self.host.is3d_host_chosen(self.h, None)
        out = np.zeros(n, dtype=np.int32)
        self.host.is3d_host_chosen(self.h, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def ptb(self):
        x, l2, z = np.empty(301), np.empty(301), np.empty(301)
        xmax = C.c_double()
        n = self.host.is3d_host_ptb(self.h, _ptr(x), _ptr(l2), _ptr(z), C.byref(xmax))
        return (x[:n], l2[:n], z[:n], xmax.value) if n else None

    def surface_column(self, k: int) -> np.ndarray:
        data = C.POINTER(C.c_double)()
        n = self.host.is3d_host_surface_column(self.h, k, C.byref(data))
        return np.ctypeslib.as_array(data, shape=(n,)).copy() if n else np.empty(0)

    def seconds(self) -> float:
        return self.host.is3d_host_seconds(self.h)

    def stats(self) -> Stats:
        st = Stats()
        self.host.is3d_host_stats(self.h, C.byref(st))
        return st


# ---- sampler conveniences (HostSession methods) ------------------------------------------------------------------
def _abi_total_yield(self):
    v = C.c_double()
    st = Stats()
    self._check(self.lib.is3d_total_yield(self.ctx, C.byref(v), C.byref(st)), "is3d_total_yield")
    return v.value, st


def _abi_cell_yields(self, n_cells: int, n_species: int, with_list: bool = True):
    tot = np.zeros(n_cells)
    lst = np.zeros((n_cells, n_species)) if with_list else None
    st = Stats()
    self._check(self.lib.is3d_cell_yields(self.ctx, _ptr(tot), _ptr(lst) if with_list else None, C.byref(st)), "is3d_cell_yields")
    return tot, lst, st


def _abi_sample(self, nevents: int, copy: bool = True, compact: bool = False):
    """Returns (structured particle array grouped by event, counts per event, stats).  copy=False returns a view of the
    library-owned pinned list plus a release callable as a 4th item (the caller must call it when done).  compact=True:
    64-byte wire records (is3d_sample_compact)."""
    plist = C.c_void_p()
    total = C.c_int64()
    counts = np.zeros(nevents, dtype=np.int64)
    st = Stats()
    fn, dt = (self.lib.is3d_sample_compact, COMPACT_DTYPE) if compact else (self.lib.is3d_sample, PARTICLE_DTYPE)
    self._check(fn(self.ctx, nevents, C.byref(plist), C.byref(total), _ptr(counts), C.byref(st)), "is3d_sample")
    n = total.value
    if plist.value and n:
        buf = (C.c_char * (n * dt.itemsize)).from_address(plist.value)
        arr = np.frombuffer(buf, dtype=dt)
        if copy:
            arr = arr.copy()
    else:
        arr = np.zeros(0, dtype=dt)
    lib = self.lib

    def release(ptr=plist):
        if ptr.value:
            lib.is3d_free_particles(ptr)
            ptr.value = None
    if copy:
        release()
        return arr, counts, st
    return arr, counts, st, release


def _abi_sample_histograms(self, ns: int, params: dict):
    g = lambda k, d: int(float(params.get(k, d)))  # noqa: E731
    yb, eb, pb, phb, tb, rb = g("y_bins", 100), g("eta_bins", 140), g("pT_bins", 100), g("phip_bins", 100), g("tau_bins", 120), g("r_bins", 60)
    h = {"dN_dy": np.zeros((ns, yb)), "dN_deta": np.zeros((ns, eb)), "dN_dphip": np.zeros((ns, phb)), "dN_pT": np.zeros((ns, pb)),
         "pT_count": np.zeros((ns, pb)), "vn_re": np.zeros((7, ns, pb)), "vn_im": np.zeros((7, ns, pb)), "dN_tau": np.zeros((ns, tb)),
         "dN_r": np.zeros((ns, rb)), "dN_phis": np.zeros((ns, phb))}
    order = ["dN_dy", "dN_deta", "dN_dphip", "dN_pT", "pT_count", "vn_re", "vn_im", "dN_tau", "dN_r", "dN_phis"]
    self._check(self.lib.is3d_sample_histograms(self.ctx, *[_ptr(h[k]) for k in order]), "is3d_sample_histograms")
    return h


HostSession.abi_total_yield = _abi_total_yield
HostSession.abi_cell_yields = _abi_cell_yields
def _abi_sample_compact(self, nevents: int, copy: bool = True):
    return _abi_sample(self, nevents, copy=copy, compact=True)


def _abi_expand(self, compact: np.ndarray) -> np.ndarray:
    """is3d_expand_particles: wire records -> full records."""
    c = np.ascontiguousarray(compact)
    out = np.zeros(len(c), dtype=PARTICLE_DTYPE)
    self._check(self.lib.is3d_expand_particles(self.ctx, _ptr(c), len(c), _ptr(out)), "is3d_expand_particles")
    return out


def _abi_sample_device(self, nevents: int):
    """(device pointer of the library-owned record array, total, counts per event, stats)"""
    plist = C.c_void_p()
    total = C.c_int64()
    counts = np.zeros(nevents, dtype=np.int64)
    st = Stats()
    self._check(self.lib.is3d_sample_device(self.ctx, nevents, C.byref(plist), C.byref(total), _ptr(counts), C.byref(st)), "is3d_sample_device")
    return plist.value, total.value, counts, st


HostSession.abi_sample = _abi_sample
HostSession.abi_sample_compact = _abi_sample_compact
HostSession.abi_expand = _abi_expand
HostSession.abi_sample_device = _abi_sample_device
HostSession.abi_sample_histograms = _abi_sample_histograms
