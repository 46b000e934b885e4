#!/bin/bash
# Front-end check after a kernel change: parity tests, latency by launch mode, launch list of the frame-to-map chain, adapter bench.
#   gpurun --timeout 600 -- 'bash tools/evidence_frontend.sh <tag>'
out=gpurun_out/${1:-fe}
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > $out/gpu_tests.log
timeout 120 python tools/f2m_modes.py > $out/f2m_modes.txt 2>&1
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_frontend.csv python tools/f2m_stamps.py > /dev/null 2>&1
if [ -x adapter/frontend_bench ]; then timeout 120 adapter/frontend_bench > $out/frontend_bench.json 2> $out/frontend_bench.err; fi
cat $out/gpu_tests.log $out/f2m_modes.txt; python tools/launch_summary.py $out/launches_frontend.csv | tail -n 30; head -c 1500 $out/frontend_bench.json
