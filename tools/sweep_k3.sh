#!/bin/bash
# K3 (famod Newton solve) build-variant sweep, run on the GPU box: tools/sweep_k3.sh "MINBLOCKS UNROLL" ...
out=gpurun_out/sweep_k3.log; : > $out
for v in "$@"; do
  set -- $v
  touch is3d2_b200/csrc/spectra_famod.cu
  python tools/build_variant.py -DIS3D_K3_MINBLOCKS=$1 -DIS3D_K3_UNROLL=$2 > /dev/null 2>&1 || { echo "B=$1 U=$2 build failed" >> $out; continue; }
  regs=$(cuobjdump -res-usage is3d2_b200/build/spectra_famod.cu.o | grep -A1 famod_setup_free | grep -o "REG:[0-9]*")
  python bench.py --df-mode 5 --steps 2 --warmup 2 --cells 200000 --no-cpu-baseline --no-sampler 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('minblocks=$1 unroll=$2 $regs', '%.4g evals/s' % d['value'], 'ms/step %.1f' % d['ms_per_step'])" >> $out
done
touch is3d2_b200/csrc/spectra_famod.cu; python -m is3d2_b200.build > /dev/null 2>&1
cat $out
