"""per-kernel medians of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py file.csv"""
import collections
import csv
import sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
for x in csv.DictReader(rows):
    k = x["Kernel Name"].split("(")[0] + " " + x["Grid Size"]
    agg.setdefault(k, []).append(int(x["Metric Value"]))
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} med={sorted(v)[len(v) // 2] / 1000:.2f} us")
