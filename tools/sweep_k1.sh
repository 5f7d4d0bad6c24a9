#!/bin/bash
# K1 launch-shape sweep (run on the GPU box): rebuilds spectra_df.cu per variant and runs a short bench.
# usage: tools/sweep_k1.sh "T B R" "T B R" ...   -> gpurun_out/sweep_k1.log
out=gpurun_out/sweep_k1.log; : > $out
for v in "$@"; do
  set -- $v
  touch is3d2_b200/csrc/spectra_df.cu
  python tools/build_variant.py -DIS3D_K1_THREADS=$1 -DIS3D_K1_MINBLOCKS=$2 -DIS3D_K1_R=$3 > /dev/null 2>&1 || { echo "T=$1 B=$2 R=$3 build failed" >> $out; continue; }
  python bench.py --steps 2 --warmup 2 --cells ${CELLS:-300000} --df-mode ${MODE:-2} --no-cpu-baseline --no-sampler 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('T=$1 B=$2 R=$3', '%.4g evals/s' % d['value'], 'kernel ms/step %.2f' % d['roofline']['kernel_ms_per_step'])" >> $out
done
touch is3d2_b200/csrc/spectra_df.cu; python -m is3d2_b200.build > /dev/null 2>&1
cat $out
