#!/bin/bash
mkdir -p gpurun_out
P="python scripts/c4_ivf_profile.py 10 1 1"
timeout 280 $P > gpurun_out/r2h_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ivf_|ir_|scan_topk|coarse|gemm_topk' -c 100 --csv --log-file gpurun_out/r2h_c4_new.csv $P > gpurun_out/r2h_ncu.log 2>&1
tail -1 gpurun_out/r2h_plain.log
AURA_IVF_OLD=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ivf_|ir_|scan_topk|coarse|gemm_topk' -c 100 --csv --log-file gpurun_out/r2h_c4_old.csv $P > gpurun_out/r2h_ncu_old.log 2>&1
P5="python scripts/c5_shard_profile.py 12500000 16384 10 64 1"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ivf_|ir_|scan_topk' -c 100 --csv --log-file gpurun_out/r2h_c5k10_new.csv $P5 > gpurun_out/r2h_ncu5.log 2>&1
