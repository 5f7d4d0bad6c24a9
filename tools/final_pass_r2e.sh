#!/bin/bash
# Last measurement pass of round 2 on ONE GPU with the final code: all GPU tests, the default bench line, the reference arm, per-mode
# and dN/dX numbers, then -- after the same commands have exited 0 without a profiler -- ncu of the two K1 launches of one pass
# (the single-class launch on its own: ncu leaves the second of two concurrent launches without metrics) and the headline json.
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/r2e_tests_final.log; cat $O/r2e_tests_final.log
python bench.py > $O/r2e_bench_final.json 2> $O/r2e_bench_final.err; echo "bench rc=$?"; cut -c1-200 $O/r2e_bench_final.json
python bench.py --impl reference > $O/r2e_bench_reference.json 2> $O/r2e_bench_reference.err; echo "reference rc=$?"; cut -c1-160 $O/r2e_bench_reference.json
tools/quick_modes.sh r2e_final
for m in 1 2 3 4; do python tools/dndx_probe.py $m 100000 2>/dev/null | tail -1; done >> $O/r2e_final_numbers.txt
python tools/polzn_probe.py 400000 2>/dev/null | tail -1 >> $O/r2e_final_numbers.txt
tail -5 $O/r2e_final_numbers.txt
N="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
$N -k regex:df_spectra_kernel -s 18 -c 2 -o $O/r2e_prof_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
$N -k regex:"df_spectra_kernel<.*4, .bool.0>" -s 9 -c 1 -o $O/r2e_prof_k1_single python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2e_prof_k1.ncu-rep | head -23 > $O/r2e_ncu_k1_pair_summary.txt; python tools/ncu_summary.py $O/r2e_prof_k1_single.ncu-rep > $O/r2e_ncu_k1_single_summary.txt
grep "fp64_cycles\|time_duration\|issue_active\|dram__bytes" $O/r2e_ncu_k1_pair_summary.txt $O/r2e_ncu_k1_single_summary.txt
python tools/make_ncu_headline.py $O/r2e_prof_k1.ncu-rep 4194304 $O/r2e_prof_k1_single.ncu-rep > /dev/null && cp profiles/ncu_k1_headline.json $O/r2e_ncu_k1_headline.json
rm -f $O/r2e_prof_k1_single.ncu-rep
