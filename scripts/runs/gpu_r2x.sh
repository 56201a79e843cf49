#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 2 32" "6 2 32" "8 2 32" "4 1 32" "8 1 32" "8 2 24"; do set -- $cfg; echo -n "rows=1000000 SAMPLE=$1 WG=$2 L=$3: "; AURA_GEMM_SAMPLE=$1 AURA_GEMM_WG=$2 AURA_SHADOW_L=$3 timeout 300 python scripts/kernel_breakdown.py 100 1000000 2>&1 | tail -1 | cut -c60-; done | tee gpurun_out/r2x2.log
