"""Instruction mix of the innermost loop of a kernel: `cuobjdump -sass -fun <mangled> obj | python tools/sass_loop.py`.
Picks the shortest backward-branch loop that contains a MUFU.RCP64H (the Cooper-Frye item loop)."""
import re
import sys
from collections import Counter

ops = []
for l in sys.stdin:
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m:
        ops.append((int(m.group(1), 16), m.group(2), l.strip()))
best = None
for a, op, l in ops:
    if op.startswith('BRA'):
        m = re.search(r'0x([0-9a-f]+)', l.split('BRA', 1)[1])
        if m:
            t = int(m.group(1), 16)
            if t < a:
                body = [o for o in ops if t <= o[0] <= a]
                if any('RCP64H' in o[1] for o in body) and (best is None or len(body) < len(best)):
                    best = body
c = Counter(o[1].split('.')[0] for o in best)
print(f"loop {best[0][0]:#x}..{best[-1][0]:#x}: {len(best)} instructions")
print(c.most_common())
fp64 = sum(v for k, v in c.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
print("FP64-pipe instructions:", fp64, " MUFU:", c.get('MUFU', 0))
if '-v' in sys.argv:
    for o in best:
        print(o[2])
