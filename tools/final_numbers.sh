#!/bin/bash
# End-of-round numbers for profiles/ (run on the GPU box): every df_mode, dN/dX, sampler variants.
out=gpurun_out/final_numbers.txt; : > $out
for m in 1 2 3 4 5; do
  python bench.py --df-mode $m --steps 2 --warmup 3 --cells 400000 --no-cpu-baseline --no-sampler 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('spectra df_mode $m: %.4g evals/s, %.1f ms/step, e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))" >> $out
done
for m in 1 2 3 4; do python tools/dndx_probe.py $m 100000 2>/dev/null | grep "^dN" | tail -1 >> $out; done
python tools/sampler_probe.py 100000 1000 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('sampler df_mode 3 fast: %.4g hadrons/s, %.1f ms, device %.1f ms, %d hadrons, acceptance %.3f' % (d['value'], 1e3*d['seconds'], d['device_ms'], d['hadrons'], d['acceptance']))" >> $out
cat $out
