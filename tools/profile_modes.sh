#!/bin/bash
# Per-mode timing + ncu launch lists (run on the GPU box).  Output under gpurun_out/.
set -x
for m in 1 3 4 5; do
  python bench.py --df-mode $m --steps 2 --warmup 3 --cells 200000 --no-cpu-baseline --no-sampler 2>/dev/null | grep '^{' > gpurun_out/bench_mode$m.json
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode2.json 2>/dev/null
# launch lists (share of the step per kernel); numbers printed under ncu are not bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_mode2.csv python bench.py --steps 1 --warmup 3 --cells 100000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_mode3.csv python bench.py --df-mode 3 --steps 1 --warmup 3 --cells 100000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_sampler.csv python tools/sampler_probe.py 100000 1000 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_mode5.csv python bench.py --df-mode 5 --steps 1 --warmup 3 --cells 100000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:df_spectra_kernel -s 3 -c 1 -o gpurun_out/prof_k1_r1c -f python bench.py --steps 1 --warmup 3 --cells 100000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:feqmod_spectra_kernel -s 3 -c 1 -o gpurun_out/prof_k2_r1 -f python bench.py --df-mode 3 --steps 1 --warmup 3 --cells 50000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:famod_setup_free_kernel -s 3 -c 1 -o gpurun_out/prof_k3_r1 -f python bench.py --df-mode 5 --steps 1 --warmup 3 --cells 50000 --no-cpu-baseline --no-sampler > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sampler_hadron_kernel -s 1 -c 1 -o gpurun_out/prof_k6_r1 -f python tools/sampler_probe.py 100000 1000 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/bench_mode*.json | cut -c1-260
