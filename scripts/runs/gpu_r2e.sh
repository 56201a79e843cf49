#!/bin/bash
mkdir -p gpurun_out
P="python scripts/c5_shard_profile.py 12500000 16384 100 64 1"
timeout 280 $P > gpurun_out/r2e_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ivf_|ir_|scan_topk' -c 300 --csv --log-file gpurun_out/r2e_launches.csv $P > gpurun_out/r2e_ncu.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/r2e_plain.log
timeout 200 python scripts/c5_shard_profile.py 12500000 16384 10 64 1 2>&1 | tail -1
timeout 200 python -m pytest tests/test_sharded_gpu.py -q -m gpu -x 2>&1 | tail -3
