#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
(time python bench.py --impl reference --steps 5 --warmup 1) 2>&1 | tail -5 | tee gpurun_out/bench_ref_n1.log | cut -c1-900
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3) 2>&1 | tail -6 | tee gpurun_out/bench_n2.log | cut -c1-1800
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1) 2>&1 | tail -5 | tee gpurun_out/bench_ref_n2.log | cut -c1-600
