"""per-opcode instruction count of the hottest loop (the backward branch whose body holds most LOP3) of a `cuobjdump -sass
-fun <kernel>` listing: python tools/sass_loop_mix.py listing.sass [pairs_per_trip]"""
import re
import sys
from collections import Counter
lines = open(sys.argv[1]).read().splitlines()
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
ins = []
for l in lines:
    m = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for a, i in ins:
    m = re.search(r'BRA\s+(?:P\d,\s*)?0x([0-9a-f]+)', i)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for (b, x) in ins if tgt <= b <= a]
        if any(x.split()[-1 if False else 0].startswith(('BAR', 'SYNCS')) or ' BAR' in x or 'SYNCS' in x for x in body):
            continue                                  # not an innermost compute loop
        n = sum('LOP3' in x for x in body)
        if best is None or n > best[0]:
            best = (n, tgt, a, body)
n, tgt, a, body = best
print(f"hot loop {hex(tgt)}..{hex(a)}: {len(body)} instructions per trip = {len(body) / pairs:.2f} per descriptor pair ({pairs:g} pairs per lane per trip)")
PIPE = {"LOP3": "alu", "VIMNMX": "alu", "VIMNMX3": "alu", "ISETP": "alu", "LEA": "alu", "VIADD": "alu", "IADD3": "alu", "SHF": "alu",
        "PRMT": "alu", "MOV": "alu", "SEL": "alu", "POPC": "xu", "IMAD": "fma", "LDS": "lsu", "STS": "lsu", "LDG": "lsu", "CREDUX": "redux",
        "REDUX": "redux", "BRA": "branch", "BAR": "branch"}
ops, pipes = Counter(), Counter()
for i in body:
    t = i.split()
    op = t[1] if t[0].startswith('@') else t[0]
    p = op.split('.')
    name = p[0] + ('.' + p[1] if p[0] == 'IMAD' and len(p) > 1 and p[1] in ('IADD', 'MOV', 'U32', 'SHL') else '')
    ops[name] += 1
    pipes[PIPE.get(p[0], "other")] += 1
for k, v in ops.most_common():
    print(f"  {k:12s} {v:4d}  {v / pairs:5.2f} / pair")
print("per pipe:", ", ".join(f"{k} {v / pairs:.2f}" for k, v in pipes.most_common()))
