"""Run only the sampler section of bench.py (for ncu launch lists / quick timing): python tools/sampler_probe.py [cells] [events]"""
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
events = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
torch.cuda.set_device(0)
os.environ["IS3D_DEVICE"] = "0"
args = types.SimpleNamespace(sampler_cells=cells, sampler_events=events, sampler_calls=3)
print(json.dumps(bench.sampler_bench(args, 0, 1, 0)))
