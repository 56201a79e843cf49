#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 32" "2 32" "1 24" "2 24"; do set -- $cfg; echo -n "WG=$1 L=$2: "; AURA_GEMM_WG=$1 AURA_SHADOW_L=$2 timeout 300 python scripts/kernel_breakdown.py 100 2>&1 | tail -1; done | tee gpurun_out/r2w.log
