#!/bin/bash
# Round-2 measurement pass on ONE GPU (run on the GPU box): headline bench line, reference arm, per-mode numbers, dN/dX and
# sampler probes, then -- each only after its command has exited 0 without a profiler -- the ncu launch list of the default bench
# command and one `ncu --set full` capture of the dominant kernel.  Everything lands in gpurun_out/.
O=gpurun_out
python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err; echo "bench rc=$?"; cut -c1-300 $O/r2_bench_final.json
python bench.py --impl reference > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err; echo "reference rc=$?"; cut -c1-300 $O/r2_bench_reference.json
for m in 1 2 3 4 5; do
  python bench.py --df-mode $m --steps 2 --warmup 3 --cells 400000 --no-cpu-baseline --no-sampler --check-cells 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('spectra df_mode $m: %.4g evals/s, %.1f ms/step, e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done > $O/r2_final_numbers.txt
for m in 1 2 3 4; do python tools/dndx_probe.py $m 100000 2>/dev/null | tail -1; done >> $O/r2_final_numbers.txt
python tools/polzn_probe.py 400000 2>/dev/null | tail -1 >> $O/r2_final_numbers.txt
cat $O/r2_final_numbers.txt
# launch list of the default bench command (cold-cache, serialised per-launch times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench_default.csv python bench.py --no-cpu-baseline > /dev/null 2>&1
python tools/launch_shares.py $O/r2_launches_bench_default.csv | head -12
# one full capture of the dominant kernels: the pair launch and the single-class launch of the first 4 194 304-cell pass of the
# timed 10 M-cell step (3 warm-up steps x 3 passes x 2 launches = 18 launches skipped)
ncu --set full --clock-control none --import-source on -f -k regex:df_spectra_kernel -s 18 -c 2 -o $O/r2_prof_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sampler --check-cells 0 > /dev/null 2>&1
python tools/ncu_summary.py $O/r2_prof_k1.ncu-rep > $O/r2_ncu_k1_summary.txt; cat $O/r2_ncu_k1_summary.txt
python tools/make_ncu_headline.py $O/r2_prof_k1.ncu-rep 4194304 > /dev/null && cp profiles/ncu_k1_headline.json $O/
ncu -i $O/r2_prof_k1.ncu-rep --page source --csv > $O/r2_k1_source.csv 2>/dev/null; wc -l $O/r2_k1_source.csv
