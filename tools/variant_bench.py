#!/usr/bin/env python
"""Time one build variant of the libraries (tools/build_variant.py output copied to variants/<name>/) with bench.py's spectra
section:  python tools/variant_bench.py variants/<name> [bench.py options]   -> one line: name, evals/s, kernel ms per step."""
import io
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from is3d2_b200 import capi  # noqa: E402

libdir = os.path.abspath(sys.argv[1])
capi.load_libraries(libdir)
import bench  # noqa: E402

sys.argv = ["bench.py", "--no-cpu-baseline", "--no-sampler", "--check-cells", "0"] + sys.argv[2:]
r, w = os.pipe()
saved = os.dup(1)
os.dup2(w, 1)
try:
    bench.main()
finally:
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(w)
data = os.read(r, 1 << 20).decode()
line = [l for l in data.splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print(f"{os.path.basename(libdir):12s} {d['value']:.4g} evals/s  kernel {d['roofline']['kernel_ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']:.4g}")
