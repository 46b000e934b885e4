#!/bin/bash
# One bounded GPU session for the headline path:  gpurun --timeout 900 -- 'bash tools/evidence_sweep.sh <tag>'
#   GPU tests, the full bench line (N = 1), the launch list of the headline part, one --set full capture of the sweep kernel.
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $out/gpu_tests.log
timeout 400 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_headline.csv \
    python bench.py --no-frontend --no-cpu-baseline --steps 5 --warmup 3 > /dev/null 2>&1
timeout 180 ncu --set full --clock-control none --import-source on -k regex:lc_sweep_kernel -s 3 -c 1 -f -o $out/lc_sweep \
    python bench.py --no-frontend --no-cpu-baseline --steps 3 --warmup 3 > $out/ncu_lc_sweep.log 2>&1
ncu -i $out/lc_sweep.ncu-rep --page raw --csv > $out/lc_sweep_raw.csv 2>/dev/null
ncu -i $out/lc_sweep.ncu-rep --page source --csv > $out/lc_sweep_source.csv 2>/dev/null
cat $out/gpu_tests.log; python tools/summ.py $out/bench_n1.json; tail -2 $out/bench_n1.err; ls -la $out
