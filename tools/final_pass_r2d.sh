#!/bin/bash
# completes tools/final_pass_r2c.sh: the single-class launch of K1 on its own (ncu leaves the second of two concurrent launches
# without metrics), the merged headline json, and the K2 pair launch at its final launch shape
O=gpurun_out
N="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
$N -k regex:df_spectra_kernel -s 18 -c 2 -o $O/r2d_prof_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
$N -k regex:"df_spectra_kernel<.*4, .bool.0>" -s 9 -c 1 -o $O/r2d_prof_k1_single python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2d_prof_k1.ncu-rep > $O/r2d_ncu_k1_summary.txt; python tools/ncu_summary.py $O/r2d_prof_k1_single.ncu-rep > $O/r2d_ncu_k1_single_summary.txt; cat $O/r2d_ncu_k1_single_summary.txt
python tools/make_ncu_headline.py $O/r2d_prof_k1.ncu-rep 4194304 $O/r2d_prof_k1_single.ncu-rep > /dev/null && cp profiles/ncu_k1_headline.json $O/r2d_ncu_k1_headline.json
$N -k regex:"feqmod_spectra_kernel<.bool.0, .*3, .bool.1>" -s 3 -c 1 -o $O/r2d_prof_k2_pair python bench.py --df-mode 3 --steps 1 --warmup 3 --cells 200000 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2d_prof_k2_pair.ncu-rep > $O/r2d_ncu_k2_pair_summary.txt; cat $O/r2d_ncu_k2_pair_summary.txt
rm -f $O/r2d_prof_k1.ncu-rep $O/r2d_prof_k1_single.ncu-rep
