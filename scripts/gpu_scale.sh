#!/bin/bash
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_n$N.log | cut -c1-700
