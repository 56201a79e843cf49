#!/bin/bash
mkdir -p gpurun_out
AURA_IVF_RB=0 timeout 600 python scripts/c5_lm_one.py 2>&1 | tail -5 | tee gpurun_out/r3b_stats.log
AURA_IVF_RB=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivf_rows_kernel --launch-skip 9 --launch-count 1 -o gpurun_out/r3b_c5_lm_heavy -f python scripts/c5_lm_one.py > gpurun_out/r3b_ncu.log 2>&1; echo "ncu rc=$?"
