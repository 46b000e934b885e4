"""per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py launches.csv"""
import csv
import sys
from collections import defaultdict
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", "")); u = r.get("Metric Unit", "ns")
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
    k = r["Kernel Name"].split("(")[0]
    tot[k][0] += 1; tot[k][1] += us
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} n={n:4d} avg={us / n:8.2f} us total={us:9.1f} us")
