"""print the key fields of bench.py JSON lines: python tools/summ.py file [file...]"""
import json
import sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e); continue
    r = d.get("roofline", {})
    print(f"{f}: value={d.get('value'):.1f} {d.get('unit')} n_gpus={d.get('n_gpus')} ms/step={d.get('ms_per_step'):.3f} "
          f"e2e={d.get('e2e', {}).get('value', 0):.1f} sweep_ms={r.get('avg_launch_ms')} frac={r.get('frac')} "
          f"mix_frac={r.get('frac_of_issued_mix')} non_sweep_us={r.get('non_sweep_us_per_step')} clocks={d.get('clocks')} ok={d.get('config', {}).get('result_ok')}")
    if "frontend" in d:
        print("   frontend:", json.dumps(d["frontend"]))
    if "cpu_baseline" in d:
        print("   cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
