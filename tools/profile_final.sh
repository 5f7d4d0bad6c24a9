#!/bin/bash
# One `ncu --set full` capture per hot kernel with the end-of-round code (run on the GPU box; every command has been run
# without a profiler before).  Reports land in gpurun_out/prof_final_*.ncu-rep; tools/ncu_summary.py extracts the metrics.
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:df_spectra_kernel -s 3 -c 1 -o gpurun_out/prof_final_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler > /dev/null 2>&1
$N -k regex:feqmod_spectra_kernel -s 3 -c 1 -o gpurun_out/prof_final_k2 python bench.py --df-mode 3 --steps 1 --warmup 3 --cells 200000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
$N -k regex:famod_setup_free_kernel -s 3 -c 1 -o gpurun_out/prof_final_k3 python bench.py --df-mode 5 --steps 1 --warmup 3 --cells 200000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
$N -k regex:dndx_df_kernel -c 1 -o gpurun_out/prof_final_k4 python tools/dndx_probe.py 2 50000 > /dev/null 2>&1
$N -k regex:sampler_hadron_kernel -s 1 -c 1 -o gpurun_out/prof_final_k6 python tools/sampler_probe.py 100000 1000 > /dev/null 2>&1
$N -k regex:polarization_kernel -s 1 -c 1 -o gpurun_out/prof_final_k7 python tools/polzn_probe.py 200000 > /dev/null 2>&1
ls -la gpurun_out/prof_final_*.ncu-rep
