#!/bin/bash
# One bounded GPU session for the headline path:  gpurun --timeout 1200 -- 'bash tools/evidence_sweep.sh <tag>'
#   GPU tests, the full bench line (N = 1), the launch list of the headline part, one --set full capture of the sweep kernel,
#   compute-sanitizer over the sweep / front-end tests (the kernels changed in round 2).
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $out/gpu_tests.log
timeout 400 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_headline.csv \
    python bench.py --no-frontend --no-cpu-baseline --steps 5 --warmup 3 > /dev/null 2>&1
timeout 180 ncu --set full --clock-control none --import-source on -k regex:lc_sweep_range -s 3 -c 1 -f -o $out/lc_sweep \
    python bench.py --no-frontend --no-cpu-baseline --steps 3 --warmup 3 > $out/ncu_lc_sweep.log 2>&1
ncu -i $out/lc_sweep.ncu-rep --page raw --csv > $out/lc_sweep_raw.csv 2>/dev/null
rm -f $out/lc_sweep.ncu-rep
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/launches_frontend.csv \
    python tools/run_frontend.py > /dev/null 2>&1
( echo "compute-sanitizer --tool memcheck pytest tests/test_gpu_sweep.py tests/test_gpu_resident_map.py tests/test_gpu_orb.py -m gpu -k 'not full_size'"
  timeout 280 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_sweep.py tests/test_gpu_resident_map.py tests/test_gpu_orb.py -m gpu -q -x -k "not full_size" 2>&1 | tail -5
  echo "compute-sanitizer --tool memcheck pytest tests/test_gpu_parity.py tests/test_gpu_ref_build.py -m gpu -k 'frame_to or ransac or guided'"
  timeout 280 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_build.py -m gpu -q -x -k "frame_to or ransac or guided" 2>&1 | tail -5
  echo "compute-sanitizer --tool racecheck pytest tests/test_gpu_sweep.py -m gpu -k 'lc_scores_vs_oracle or topk'"
  timeout 280 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_sweep.py -m gpu -q -x -k "lc_scores_vs_oracle or topk" 2>&1 | tail -5
) > $out/sanitizer.txt 2>&1
cat $out/gpu_tests.log; python tools/summ.py $out/bench_n1.json | cut -c1-900; tail -n 2 $out/bench_n1.err; cat $out/sanitizer.txt; ls -la $out
