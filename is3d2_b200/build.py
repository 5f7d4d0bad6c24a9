"""Build the in-tree native libraries:
  is3d2_b200/libis3d_b200.so   CUDA kernels + C ABI (nvcc, sm_100a only)
  is3d2_b200/libis3d_host.so   C++ host layer (readers, tables, EmissionFunctionArray, IS3D)
  is3d2_b200/iS3D_b200.e       drop-in executable
Built artefacts are git-ignored but travel to the GPU box with the snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
CU_SOURCES = ["api.cu", "spectra_df.cu", "spectra_feqmod.cu", "dndx.cu", "sampler.cu", "spectra_famod.cu", "polarization.cu",
              "fp64_peak.cu", "comm.cu"]
HOST_SOURCES = ["io.cpp", "surface.cpp", "pdg.cpp", "deltaf.cpp", "emission.cpp", "is3d.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--use_fast_math=false"]
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd: list[str]) -> None:
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force: bool = False, verbose_ptxas: bool = False, variant_flags: list[str] | None = None) -> None:
    """variant_flags: extra -D defines for the launch-shape sweeps under tools/ (tools/build_variant.py); the product build
    takes none and reads no flags from the environment."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    lib = os.path.join(HERE, "libis3d_b200.so")
    cu = [os.path.join(CSRC, f) for f in CU_SOURCES]
    deps = cu + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
        [os.path.join(HERE, "..", "include", "is3d_b200.h")]
    if force or variant_flags or _newer(lib, deps):
        force = force or bool(variant_flags)
        objs = []
        os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
        procs = []
        for src in cu:
            obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
            objs.append(obj)
            if force or _newer(obj, deps):
                cmd = [nvcc] + [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + \
                      list(variant_flags or []) + \
                      (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-c", src, "-o", obj]
                print(" ".join(cmd), flush=True)
                procs.append(subprocess.Popen(cmd))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("nvcc failed")
        _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs + ["-ldl"])
    hostlib = os.path.join(HERE, "libis3d_host.so")
    hs = [os.path.join(HOST, f) for f in HOST_SOURCES]
    hdeps = hs + [os.path.join(HOST, "is3d_host.hpp"), os.path.join(HERE, "..", "include", "is3d_host.h"), lib] + \
        [os.path.join(CSRC, f) for f in ("dftables.cuh", "gauss_thermal.cuh", "common.cuh")]
    if force or _newer(hostlib, hdeps):
        _run([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-pthread", "-o", hostlib] + hs +
             ["-L" + HERE, "-lis3d_b200", "-Wl,-rpath,$ORIGIN"])
    exe = os.path.join(HERE, "iS3D_b200.e")
    if force or _newer(exe, [os.path.join(HOST, "main.cpp"), hostlib]):
        _run([CXX, "-std=c++17", "-O2", "-pthread", "-o", exe, os.path.join(HOST, "main.cpp"), "-L" + HERE, "-lis3d_host",
              "-lis3d_b200", "-Wl,-rpath,$ORIGIN"])


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
