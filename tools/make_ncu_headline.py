#!/usr/bin/env python
"""profiles/ncu_k1_headline.json from one `ncu --set full` capture of the headline kernels of ONE pass (the pair launch and the
single-class launch of df_spectra_kernel):
    python tools/make_ncu_headline.py report.ncu-rep cells_of_the_captured_pass [second_report.ncu-rep ...]
(ncu leaves the metrics of the second of two concurrent launches empty: capture the single-class launch on its own and pass that
report as well -- for every kernel the first record with real numbers is taken)
bench.py reports roofline.traffic from this file ONLY when the inner-loop SASS hashes recorded here equal those of the library
it has loaded (is3d2_b200/sassinfo.py), i.e. when the capture is of the same kernel code."""
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from is3d2_b200 import sassinfo  # noqa: E402

reps, cells = [sys.argv[1]] + sys.argv[3:], int(sys.argv[2])
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}


by_name = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def num(vals, name):
        return float(vals[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1.0)

    for vals in rows[2:]:
        if len(vals) < len(hdr):
            continue
        k = {"kernel": vals[col["Kernel Name"]][:120], "gpu_time_ms": num(vals, "gpu__time_duration.sum"),
             "grid": vals[col["launch__grid_size"]], "registers": vals[col["launch__registers_per_thread"]],
             "dram_bytes_read": num(vals, "dram__bytes_read.sum"), "dram_bytes_write": num(vals, "dram__bytes_write.sum"),
             "fp64_pipe_active_pct": num(vals, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
             "issue_active_pct": num(vals, "smsp__issue_active.avg.pct_of_peak_sustained_active"), "report": os.path.basename(rep)}
        old = by_name.get(k["kernel"])
        if old is None or (old["dram_bytes_read"] != old["dram_bytes_read"] and k["dram_bytes_read"] == k["dram_bytes_read"]):
            by_name[k["kernel"]] = k
kernels = list(by_name.values())
info = sassinfo.library_info()
labels = ["df_spectra_kernel<2,1,0,0,4,0>", "df_spectra_kernel<2,1,0,0,4,1>"]
t = sum(k["gpu_time_ms"] for k in kernels)
out = {"kernels": kernels, "cells_per_launch": cells, "gpu_time_ms": t,
       "dram_bytes_read": sum(k["dram_bytes_read"] for k in kernels), "dram_bytes_write": sum(k["dram_bytes_write"] for k in kernels),
       "fp64_pipe_active_pct": sum(k["fp64_pipe_active_pct"] * k["gpu_time_ms"] for k in kernels) / t,
       "issue_active_pct": sum(k["issue_active_pct"] * k["gpu_time_ms"] for k in kernels) / t,
       "listing_sha256": {l: info["kernels"][l]["listing_sha256"] for l in labels if l in info["kernels"]},
       "library_sha256": info["library_sha256"],
       "source": "one `ncu --set full --clock-control none` capture of the two launches of one pass (ncu serialises them), tools/final_pass_r2.sh"}
path = os.path.join(REPO, "profiles", "ncu_k1_headline.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
