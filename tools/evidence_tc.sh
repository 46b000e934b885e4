#!/bin/bash
# Tensor-core sweep evidence (1 GPU):  gpurun --timeout 900 -- 'bash tools/evidence_tc.sh <tag>'
tag=${1:-tc}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8 > $out/gpu_tests.log
timeout 60 bench/tc_probe > $out/tc_probe.jsonl 2>&1
timeout 400 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_headline.csv \
    python bench.py --no-frontend --no-cpu-baseline --steps 5 --warmup 3 > /dev/null 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:lc_tc_sweep -s 3 -c 1 -f -o $out/lc_tc_sweep \
    python bench.py --no-frontend --no-cpu-baseline --steps 3 --warmup 3 > $out/ncu_lc_tc_sweep.log 2>&1
ncu -i $out/lc_tc_sweep.ncu-rep --page raw --csv > $out/lc_tc_sweep_raw.csv 2>/dev/null
ncu -i $out/lc_tc_sweep.ncu-rep --page details --csv > $out/lc_tc_sweep_details.csv 2>/dev/null
rm -f $out/lc_tc_sweep.ncu-rep
( echo "compute-sanitizer --tool memcheck pytest tests/test_gpu_sweep.py tests/test_gpu_fuzz.py -m gpu -k 'not full_size'"
  timeout 280 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fuzz.py -m gpu -q -x -k "not full_size" 2>&1 | tail -n 5
) > $out/sanitizer.txt 2>&1
cat $out/gpu_tests.log; python tools/summ.py $out/bench_n1.json | cut -c1-900; tail -n 2 $out/bench_n1.err; cat $out/sanitizer.txt; python tools/launch_summary.py $out/launches_headline.csv | head -n 6; ls -la $out
