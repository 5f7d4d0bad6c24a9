"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python tools/launch_shares.py file.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'<unnamed>::', '', name)[:80]
    v = float(row['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(row['Metric Unit'], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:11.3f} ms {n:5d} launches {100 * t / tot:6.2f} %  {k}")
