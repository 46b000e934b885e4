#!/bin/bash
# One bounded GPU session that refreshes every piece of measured evidence (run under gpurun from the repo root):
#   gpurun --timeout 420 -- 'bash tools/profile_next.sh r2'
# Writes into gpurun_out/ (copy what should be judged into profiles/).  Budget: ~4-5 minutes of box time.
tag=${1:-next}
out=gpurun_out
mkdir -p $out
# 1. parity first: the whole GPU suite
timeout 170 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/gpu_tests_$tag.log
# 2. headline bench line (N = 1)
timeout 150 python bench.py > $out/bench_${tag}_n1.json 2> $out/bench_${tag}_n1.err
# 3. launch lists: KLT tracking (v2 kernels have only been timed end to end), ORB, front end
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_klt_$tag.csv python tools/run_klt.py --once > /dev/null 2>&1
# 4. full captures of the two kernels without one: klt_track_kernel (second launch), klt_prune_kernel
timeout 60 ncu --set full --clock-control none --import-source on -k regex:klt_track -s 1 -c 1 -f -o $out/klt_track_$tag python tools/run_klt.py --once > $out/ncu_klt_track_$tag.log 2>&1
timeout 40 ncu --set full --clock-control none --import-source on -k regex:klt_prune -s 1 -c 1 -f -o $out/klt_prune_$tag python tools/run_klt.py --once > $out/ncu_klt_prune_$tag.log 2>&1
# 5. end-to-end KLT numbers
timeout 40 python tools/run_klt.py > $out/klt_timing_$tag.json 2> $out/klt_timing_$tag.err
tail -3 $out/gpu_tests_$tag.log; head -c 600 $out/bench_${tag}_n1.json; echo; cat $out/klt_timing_$tag.json
