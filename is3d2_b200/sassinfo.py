"""Instruction mix of a kernel's inner loop, read from the SASS of the shared library that is actually loaded.

bench.py uses it to state the EXECUTED FP64-pipe work of the dominant kernel (FP64 instructions per class-evaluation x 2
flops) instead of a number copied from an old profile: the figures are derived from `cuobjdump -sass` of the very
libis3d_b200.so the process loaded and cached beside it, keyed by the library's SHA-256.
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
from collections import Counter

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libis3d_b200.so")
CACHE = os.path.join(HERE, "libis3d_b200.sass.json")
FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")

# kernels the benchmark can name: label -> (substring of the mangled name, class-evaluations per trip of the inner loop)
# template arguments: <MODE, BARYON, REGULATE, OUTFLOW, R, PAIR>; a pair slot is two class-evaluations from one exponential
KERNELS = {
    "df_spectra_kernel<1,1,0,0,4,0>": ("df_spectra_kernelILi1ELb1ELb0ELb0ELi4ELb0EE", 4),
    "df_spectra_kernel<2,1,0,0,4,0>": ("df_spectra_kernelILi2ELb1ELb0ELb0ELi4ELb0EE", 4),
    "df_spectra_kernel<1,1,0,0,4,1>": ("df_spectra_kernelILi1ELb1ELb0ELb0ELi4ELb1EE", 8),
    "df_spectra_kernel<2,1,0,0,4,1>": ("df_spectra_kernelILi2ELb1ELb0ELb0ELi4ELb1EE", 8),
    "df_spectra_kernel<1,0,0,0,4,0>": ("df_spectra_kernelILi1ELb0ELb0ELb0ELi4ELb0EE", 4),
    "df_spectra_kernel<2,0,0,0,4,0>": ("df_spectra_kernelILi2ELb0ELb0ELb0ELi4ELb0EE", 4),
}


def sha256_of(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def _loop_mix(lines: list[str]) -> dict | None:
    """Shortest backward-branch loop containing a MUFU.RCP64H (the Cooper-Frye item loop)."""
    ops = []
    for l in lines:
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops.append((int(m.group(1), 16), m.group(2), l))
    best = None
    for a, op, l in ops:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", l.split("BRA", 1)[1])
            if m:
                t = int(m.group(1), 16)
                if t < a:
                    body = [o for o in ops if t <= o[0] <= a]
                    if any("RCP64H" in o[1] for o in body) and (best is None or len(body) < len(best)):
                        best = body
    if best is None:
        return None
    c = Counter(o[1].split(".")[0] for o in best)
    listing = "\n".join(re.sub(r"\s+", " ", o[2].split("*/", 1)[1].split("/*")[0]).strip() for o in best)
    return {"instructions": len(best), "fp64": sum(v for k, v in c.items() if k in FP64_OPS), "mufu": c.get("MUFU", 0),
            "mix": dict(c.most_common()), "listing_sha256": hashlib.sha256(listing.encode()).hexdigest(), "listing": listing}


def scan_library(lib: str = LIB, cuobjdump: str | None = None) -> dict:
    """{label: loop mix} for every kernel of KERNELS found in `lib` (one full `cuobjdump -sass` pass, ~10 s)."""
    exe = cuobjdump or os.environ.get("CUOBJDUMP") or "/usr/local/cuda/bin/cuobjdump"
    p = subprocess.Popen([exe, "-sass", lib], stdout=subprocess.PIPE, text=True, errors="replace")
    want = {label: pat for label, (pat, _) in KERNELS.items()}
    out, cur, buf = {}, None, []

    def flush():
        if cur is not None:
            mix = _loop_mix(buf)
            if mix is not None:
                mix["evals_per_trip"] = KERNELS[cur][1]
                out[cur] = mix

    for line in p.stdout:
        if "Function :" in line:
            flush()
            cur, buf = None, []
            for label, pat in want.items():
                if pat in line:
                    cur = label
        elif cur is not None:
            buf.append(line)
    flush()
    if p.wait() != 0:
        raise RuntimeError("cuobjdump failed")
    return out


def library_info(lib: str = LIB, refresh: bool = False) -> dict:
    """Cached scan of `lib`; the cache is trusted only when it records the SHA-256 of the file on disk."""
    sha = sha256_of(lib)
    if not refresh and os.path.exists(CACHE):
        try:
            c = json.load(open(CACHE))
            if c.get("library_sha256") == sha:
                return c
        except (OSError, ValueError):
            pass
    info = {"library_sha256": sha, "kernels": scan_library(lib)}
    try:
        with open(CACHE, "w") as f:
            json.dump(info, f, indent=1)
    except OSError:
        pass
    return info


if __name__ == "__main__":
    import sys
    info = library_info(refresh="--refresh" in sys.argv)
    for label, mix in info["kernels"].items():
        print(f"{label}: {mix['instructions']} instructions, {mix['fp64']} FP64-pipe, {mix['mufu']} MUFU per {mix['evals_per_trip']} "
              f"class-evaluations  [{mix['listing_sha256'][:12]}]")
        if "-v" in sys.argv:
            print(mix["listing"])
