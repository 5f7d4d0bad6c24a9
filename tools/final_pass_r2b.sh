#!/bin/bash
# Round-2 ncu captures that complete tools/final_pass_r2.sh: the single-class launch of K1 on its own, and the other kernels
# at small sizes (each command has run without a profiler in final_pass_r2.sh).
O=gpurun_out
N="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
$N -k regex:"df_spectra_kernel<.*4, .bool.0>" -s 9 -c 1 -o $O/r2_prof_k1_single python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2_prof_k1_single.ncu-rep > $O/r2_ncu_k1_single_summary.txt; cat $O/r2_ncu_k1_single_summary.txt
$N -k regex:"feqmod_spectra_kernel<.bool.0, .*4, .bool.1>" -s 3 -c 1 -o $O/r2_prof_k2_pair python bench.py --df-mode 3 --steps 1 --warmup 3 --cells 200000 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2_prof_k2_pair.ncu-rep > $O/r2_ncu_k2_pair_summary.txt; cat $O/r2_ncu_k2_pair_summary.txt
$N -k regex:famod_setup_free_kernel -s 3 -c 1 -o $O/r2_prof_k3 python bench.py --df-mode 5 --steps 1 --warmup 3 --cells 200000 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2_prof_k3.ncu-rep > $O/r2_ncu_k3_summary.txt; cat $O/r2_ncu_k3_summary.txt
$N -k regex:"dndx_df_kernel<.*.bool.1>" -c 1 -o $O/r2_prof_k4_pair python tools/dndx_probe.py 2 50000 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2_prof_k4_pair.ncu-rep > $O/r2_ncu_k4_pair_summary.txt; cat $O/r2_ncu_k4_pair_summary.txt
$N -k regex:sampler_hadron_kernel -s 6 -c 1 -o $O/r2_prof_k6 python tools/sampler_probe.py > /dev/null 2>&1
python tools/ncu_summary.py $O/r2_prof_k6.ncu-rep > $O/r2_ncu_k6_summary.txt; cat $O/r2_ncu_k6_summary.txt
rm -f $O/r2_prof_k1.ncu-rep
