#!/usr/bin/env python
"""Rebuild libis3d_b200.so with launch-shape defines for the sweeps recorded under profiles/ (not a product path):
    python tools/build_variant.py -DIS3D_K1_THREADS=256 -DIS3D_K1_MINBLOCKS=2 -DIS3D_K1_R=4
Only -D flags are accepted.  Run `python -m is3d2_b200.build --force` afterwards to restore the product build."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from is3d2_b200 import build  # noqa: E402

flags = sys.argv[1:]
bad = [f for f in flags if not f.startswith("-D")]
if bad:
    raise SystemExit(f"only -D defines are accepted, got {bad}")
build.build(force=True, variant_flags=flags)
