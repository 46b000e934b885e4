#!/bin/bash
# multi-GPU session: gpurun --gpus N --timeout 900 -- 'bash tools/multi_gpu_session.sh N tag'
n=${1:-2}; tag=${2:-mg}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -8 > $out/multirank_test.log
PSLAM_LC_P2P=0 timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -4 > $out/multirank_test_nccl.log
run() { # gpus extra-env label
  env $2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $1 --steps 20 --warmup 3 $4 > $out/bench_$3.json 2> $out/bench_$3.err
}
run $n "PSLAM_X=1" n${n}_p2p ""
run $n "PSLAM_LC_P2P=0" n${n}_nccl "--no-cpu-baseline"
if [ "$n" -ge 4 ]; then run 2 "PSLAM_X=1" n2_p2p "--no-cpu-baseline"; run 4 "PSLAM_X=1" n4_p2p "--no-cpu-baseline"; fi
timeout 300 python bench.py --no-frontend --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/multirank_test.log $out/multirank_test_nccl.log; python tools/summ.py $out/bench_*.json; for f in $out/*.err; do tail -n 3 $f; done
