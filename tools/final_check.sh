#!/bin/bash
# Last pass of the session on the GPU box: GPU test-suite, smoke(), both bench arms, per-mode numbers, 10 M-cell run.
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$? lines=$(wc -l < $O/bench_final.json)"; cut -c1-300 $O/bench_final.json
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"; cut -c1-200 $O/bench_reference.json
bash tools/final_numbers.sh > /dev/null 2>&1; cat $O/final_numbers.txt
python bench.py --cells 10000000 --steps 2 --warmup 1 --no-cpu-baseline --no-sampler 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('10 M cells on one GPU: %.4g evals/s, %.1f ms/step, e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))" | tee $O/bench_10M.txt
