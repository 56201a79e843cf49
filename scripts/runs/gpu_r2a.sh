#!/bin/bash
# round 2, call a: launch breakdown of one C5 shard batch (k=100, nprobe 64) before the one-pass selection work
mkdir -p gpurun_out
P="python scripts/c5_shard_profile.py 12500000 16384 100 64 2"
timeout 280 $P > gpurun_out/r2a_plain.log 2>gpurun_out/r2a_plain_err.log &&
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ivf_|gemm_topk|cand_|ib_|normalize|coarse|centroid_terms|scan_topk' --csv --log-file gpurun_out/r2a_launches.csv $P > gpurun_out/r2a_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2a_plain.log; tail -3 gpurun_out/r2a_plain_err.log
P2="python scripts/c5_shard_profile.py 12500000 16384 10 64 2"
timeout 280 $P2 > gpurun_out/r2a_k10.log 2>&1; tail -1 gpurun_out/r2a_k10.log
