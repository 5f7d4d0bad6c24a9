#!/bin/bash
# Round-2 (second session) measurement pass on ONE GPU: headline bench line, reference arm, the margin-off line, per-mode numbers,
# then -- each only after its command has exited 0 without a profiler -- the ncu launch list of the default bench command and
# `ncu --set full` captures of the dominant kernels.  Everything lands in gpurun_out/.
O=gpurun_out
python bench.py > $O/r2c_bench_final.json 2> $O/r2c_bench_final.err; echo "bench rc=$?"; cut -c1-200 $O/r2c_bench_final.json
python bench.py --impl reference > $O/r2c_bench_reference.json 2> $O/r2c_bench_reference.err; echo "reference rc=$?"; cut -c1-200 $O/r2c_bench_reference.json
python bench.py --negligible-margin 0 --no-cpu-baseline --no-sampler > $O/r2c_bench_margin_off.json 2> /dev/null; cut -c1-200 $O/r2c_bench_margin_off.json
tools/quick_modes.sh r2c_final
python tools/polzn_probe.py 400000 2>/dev/null | tail -1 >> $O/r2c_final_numbers.txt
# launch list of the default bench command (cold-cache, serialised per-launch times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2c_launches_bench_default.csv python bench.py --no-cpu-baseline > /dev/null 2>&1
python tools/launch_shares.py $O/r2c_launches_bench_default.csv | head -14 | tee $O/r2c_launch_shares.txt
# the pair launch and the single-class launch of the first 4 194 304-cell pass of the timed 10 M-cell step
# (3 warm-up steps x 3 passes x 2 launches = 18 launches skipped)
N="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
$N -k regex:df_spectra_kernel -s 18 -c 2 -o $O/r2c_prof_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2c_prof_k1.ncu-rep > $O/r2c_ncu_k1_summary.txt; cat $O/r2c_ncu_k1_summary.txt
python tools/make_ncu_headline.py $O/r2c_prof_k1.ncu-rep 4194304 > /dev/null && cp profiles/ncu_k1_headline.json $O/r2c_ncu_k1_headline.json
ncu -i $O/r2c_prof_k1.ncu-rep --page source --csv > $O/r2c_k1_source.csv 2>/dev/null; wc -l $O/r2c_k1_source.csv
$N -k regex:"feqmod_spectra_kernel<.bool.0, .*4, .bool.1>" -s 3 -c 1 -o $O/r2c_prof_k2_pair python bench.py --df-mode 3 --steps 1 --warmup 3 --cells 200000 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2c_prof_k2_pair.ncu-rep > $O/r2c_ncu_k2_pair_summary.txt; cat $O/r2c_ncu_k2_pair_summary.txt
