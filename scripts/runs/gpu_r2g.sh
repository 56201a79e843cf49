#!/bin/bash
mkdir -p gpurun_out
P="python scripts/c5_shard_profile.py 12500000 16384 100 64 1"
timeout 280 $P > gpurun_out/r2g_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivf_rows_kernel -c 3 -o gpurun_out/r2g_c5_rows $P > gpurun_out/r2g_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2g_ncu.log
