"""is3d2_b200: B200-native Cooper-Frye particlization hot path (drop-in for iS3D2's EmissionFunctionArray compute
members).  The product is the native code under is3d2_b200/csrc (CUDA, C ABI) and is3d2_b200/host (C++); this
package is a thin ctypes binding used by the tests and bench.py.  There is no CPU fallback: the compute entry
points raise if libis3d_b200.so is missing or no sm_100 GPU is present."""
from .capi import (HostSession, Is3dError, Params, Stats, lib, host_lib, load_libraries, SURFACE_COLUMNS)  # noqa: F401
