"""CPU test: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--n-kf", "48"], capture_output=True, text=True, timeout=600, cwd=ROOT)   # a small map: the CPU suite stays short
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "Gcmp/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    assert line["config"]["whole_map_per_step"] is True and line["config"]["keyframes_per_step"] == 48
    assert "cv2_gcmps" in line["cpu_baseline"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_bench_lines_carry_the_contract():
    """the bench lines committed under profiles/ (measured on B200) hold every key of the contract, with sane values"""
    for name in ("bench_r1_final_n1.json", "bench_r1_late_n1.json", "bench_r1_final_n8.json", "bench_r2_n1.json", "bench_r2_n8.json",
                 "bench_r2_tc_full_n1.json", "bench_r2_tc_n8.json"):
        p = os.path.join(ROOT, "profiles", name)
        line = json.loads(open(p).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert key in line, (name, key)
        assert line["unit"] == "Gcmp/s" and line["value"] > 500 * line["n_gpus"] and line["higher_is_better"] is True
        assert line["warmup"] >= 3 and line["gpu_launches"] > 0 and line["vs_baseline"] is None
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0 and line["e2e"]["value"] > 0
        r = line["roofline"]
        assert r["bound"] in ("hbm", "tensor", "int-pipe") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert "workload" in line["config"] and "model" not in line["config"]
        if name.startswith("bench_r2"):      # round 2: the result of every timed configuration is checked against the oracle
            assert line["config"]["result_ok"] is True
        if "_tc_" in name:                   # tensor-core form: tensor roofline, int8 ops
            assert r["bound"] == "tensor" and r["unit"] == "TOP/s" and line["value"] > 2500 * line["n_gpus"] * 0.8
            assert line["config"]["sweep_form"].startswith("tensor") and line["config"]["tensor_wait_timeouts"] == 0
        if line["n_gpus"] == 1 and "cpu_baseline" in line:
            assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    late = json.loads(open(os.path.join(ROOT, "profiles", "bench_r1_late_n1.json")).read().strip().splitlines()[-1])
    fe = late["frontend"]
    assert fe["orb_describe"]["bit_exact_vs_cv2"] is True and fe["orb_describe"]["detect"]["bit_exact_vs_cv2_order_included"] is True
    assert fe["native_cpp"]["resident_equals_host_map"] is True and fe["native_cpp"]["frame_to_map_ms"] < 1.0   # north_star: < 1 ms
    assert fe["c2_sequence"]["parity_bit_exact"] is True
