#!/bin/bash
# K2 (feqmod spectra kernel) species-per-thread sweep on the GPU box: rebuilds spectra_feqmod.cu per variant and runs a short bench
# for df_mode 3, 4, 5.   usage: tools/sweep_k2.sh R R ...   -> gpurun_out/sweep_k2.log
out=gpurun_out/sweep_k2.log; : > $out
for R in "$@"; do
  touch is3d2_b200/csrc/spectra_feqmod.cu
  python tools/build_variant.py -DIS3D_K2_R=$R > /dev/null 2>&1 || { echo "R=$R build failed" >> $out; continue; }
  for M in ${MODES:-3 4 5}; do
    python bench.py --steps 2 --warmup 2 --cells ${CELLS:-300000} --df-mode $M --no-cpu-baseline --no-sampler 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('R=$R df_mode=$M', '%.4g evals/s' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'spectra kernel ms/step %.2f' % d['roofline']['kernel_ms_per_step'])" >> $out
  done
done
touch is3d2_b200/csrc/spectra_feqmod.cu; python -m is3d2_b200.build > /dev/null 2>&1
cat $out
