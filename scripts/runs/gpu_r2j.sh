#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --legs none --no-graph"
timeout 200 $B > gpurun_out/r2j_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_topk|normalize|scan_topk|f32_to_bf16' -c 60 --csv --log-file gpurun_out/r2j_launches.csv $B > gpurun_out/r2j_ncu.log 2>&1
echo rc=$?
timeout 100 python -m pytest tests/test_tensorcore_gpu.py -q -m gpu -x -k shadow 2>&1 | tail -3
