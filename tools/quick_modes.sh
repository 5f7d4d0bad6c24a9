#!/bin/bash
# quick per-mode spectra numbers (400 k cells) + dN/dX probes: tools/quick_modes.sh <tag>
O=gpurun_out; T=${1:-quick}
for m in 1 2 3 4 5; do
  python bench.py --df-mode $m --steps 2 --warmup 3 --cells 400000 --no-cpu-baseline --no-sampler --check-cells 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('spectra df_mode $m: %.4g evals/s, %.1f ms/step, e2e %.4g, frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'] or 0.0))"
done > $O/${T}_numbers.txt
for m in 1 2 3 4; do python tools/dndx_probe.py $m 100000 2>/dev/null | tail -1; done >> $O/${T}_numbers.txt
cat $O/${T}_numbers.txt
