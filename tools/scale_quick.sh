#!/bin/bash
# quick scaling lines on one multi-GPU box: gpurun --gpus 8 --timeout 400 -- 'bash tools/scale_quick.sh tag'
out=gpurun_out/${1:-sq}; mkdir -p $out
run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $1 --steps 20 --warmup 3 $2 > $out/bench_n$1.json 2> $out/bench_n$1.err; }
run 8 ""
run 4 "--no-cpu-baseline"
run 2 "--no-cpu-baseline"
python tools/summ.py $out/bench_n8.json $out/bench_n4.json $out/bench_n2.json | cut -c1-230
