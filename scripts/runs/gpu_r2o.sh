#!/bin/bash
mkdir -p gpurun_out
for R in 125000 250000 1000000; do timeout 120 python scripts/c2_shard_profile.py $R 30 2>&1 | tail -1; done
P="python scripts/c2_shard_profile.py 125000 3"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_topk|normalize|pack|merge' -c 40 --csv --log-file gpurun_out/r2o_launches.csv $P > gpurun_out/r2o_ncu.log 2>&1
