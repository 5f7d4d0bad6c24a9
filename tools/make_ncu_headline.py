#!/usr/bin/env python
"""profiles/ncu_k1_headline.json from one `ncu --set full` capture of the headline kernel:
    python tools/make_ncu_headline.py report.ncu-rep cells_of_the_captured_launch
bench.py reports roofline.traffic from this file ONLY when the inner-loop SASS hash recorded here equals the one of the library
it has loaded (is3d2_b200/sassinfo.py), i.e. when the capture is of the same kernel code."""
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from is3d2_b200 import sassinfo  # noqa: E402

rep, cells = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def num(name):
    v = float(vals[col[name]].replace(",", ""))
    u = units[col[name]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}.get(u, 1.0)


kernel = vals[col["Kernel Name"]]
label = "df_spectra_kernel<2,1,0,0,4,0>"
mix = sassinfo.library_info()["kernels"][label]
out = {"kernel": kernel, "label": label, "cells_per_launch": cells, "gpu_time_ms": num("gpu__time_duration.sum"),
       "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
       "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "listing_sha256": mix["listing_sha256"], "inner_loop_instructions": mix["instructions"], "inner_loop_fp64": mix["fp64"],
       "library_sha256": sassinfo.library_info()["library_sha256"],
       "source": "one `ncu --set full --clock-control none` capture, tools/final_pass_r2.sh"}
path = os.path.join(REPO, "profiles", "ncu_k1_headline.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
