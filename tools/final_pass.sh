#!/bin/bash
# End-of-session measurement pass (run on the GPU box): headline bench line, reference arm, ncu launch list of the bench
# command, one `ncu --set full` capture per changed kernel (each after its command has run without a profiler), and the
# per-mode numbers.  Everything lands in gpurun_out/.
O=gpurun_out
python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"; cut -c1-400 $O/bench_final.json
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"; cut -c1-300 $O/bench_reference.json
bash tools/final_numbers.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:df_spectra_kernel -s 3 -c 1 -o $O/prof_u_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler > /dev/null 2>&1
$N -k regex:feqmod_spectra_kernel -s 3 -c 1 -o $O/prof_u_k2 python bench.py --df-mode 3 --steps 1 --warmup 3 --cells 200000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
$N -k regex:dndx_df_kernel -c 1 -o $O/prof_u_k4 python tools/dndx_probe.py 2 50000 > /dev/null 2>&1
$N -k regex:dndx_feqmod_kernel -c 1 -o $O/prof_u_k4m3 python tools/dndx_probe.py 3 50000 > /dev/null 2>&1
ls -la $O/*.ncu-rep
