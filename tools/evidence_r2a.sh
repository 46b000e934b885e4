#!/bin/bash
# Round-2 evidence session A (bounded): GPU tests, bench line, sanitizer over K9-K13, --set full captures of the kernels
# round 1 left without one.   gpurun --timeout 900 -- 'bash tools/evidence_r2a.sh'
out=gpurun_out/r2a
mkdir -p $out
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/gpu_tests.log
timeout 200 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
# sanitizer over the kernels added late in round 1 (KLT, uncertainty, ORB incl. the small-image fresh-context case)
( echo "compute-sanitizer --tool memcheck pytest tests/test_gpu_klt.py tests/test_gpu_uncertainty.py tests/test_gpu_orb.py -m gpu"
  timeout 280 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_klt.py tests/test_gpu_uncertainty.py tests/test_gpu_orb.py -m gpu -q -x 2>&1 | tail -6
  echo "compute-sanitizer --tool racecheck pytest tests/test_gpu_klt.py tests/test_gpu_uncertainty.py -m gpu -k 'not fuzz'"
  timeout 280 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_klt.py tests/test_gpu_uncertainty.py -m gpu -q -x -k "not fuzz" 2>&1 | tail -6
) > $out/sanitizer.txt 2>&1
cap() {  # name regex skip script
  timeout 90 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $out/$1 python $4 > $out/ncu_$1.log 2>&1
  ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2>/dev/null
}
cap klt_track klt_track 1 "tools/run_klt.py --once"
cap klt_prune klt_prune 1 "tools/run_klt.py --once"
cap unc_batch uncertainty_batch 1 tools/run_uncertainty.py
cap orb_fast orb_fast_score 1 tools/run_orb.py
cap orb_cand orb_candidates 1 tools/run_orb.py
cap orb_harris orb_harris_angle 1 tools/run_orb.py
cap orb_describe orb_describe 1 tools/run_orb.py
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_klt.csv python tools/run_klt.py --once > /dev/null 2>&1
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_unc.csv python tools/run_uncertainty.py > /dev/null 2>&1
rm -f $out/orb_*.ncu-rep   # keep the raw csv only (size)
cat $out/gpu_tests.log; head -c 400 $out/bench_n1.json; echo; cat $out/sanitizer.txt; ls -la $out
