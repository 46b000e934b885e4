"""key metrics of `ncu -i x.ncu-rep --page raw --csv` exports -> one JSON: python tools/ncu_raw_summary.py out.json a_raw.csv ..."""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
out = {}
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        out[path] = "empty"; continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")]}
        for k in KEYS:
            if k in h:
                i = h.index(k)
                try:
                    d[k] = [float(r[i]), units[i]]
                except ValueError:
                    d[k] = [r[i], units[i]]
        stalls = {h[i].split("smsp__average_warp")[-1]: float(r[i]) for i in range(len(h))
                  if h[i].startswith("smsp__average_warps_issue_stalled_") and h[i].endswith("_per_issue_active.ratio") and r[i]}
        d["top_stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        out.setdefault(path.split("/")[-1].replace("_raw.csv", ""), []).append(d)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print("wrote", sys.argv[1], list(out))
