"""hottest SASS lines of an `ncu --page source --csv` export: python tools/ncu_hot.py file.csv [top]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(h)]
ci = {n: h.index(n) for n in ("Source", "# Samples", "Instructions Executed")}
stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body), "executed", sum(int(r[ci["Instructions Executed"]] or 0) for r in body))
agg = {}
for r in body:
    for i in stall:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda k: -int(body[k][ci["# Samples"]] or 0))[:top]
for k in sorted(order):
    r = body[k]
    st = sorted(((int(r[i] or 0), h[i]) for i in stall), reverse=True)[:2]
    print(f"{k:5d} {int(r[ci['# Samples']]):6d} {r[ci['Instructions Executed']]:>8s}  {r[ci['Source']].strip()[:70]:70s} {st}")
