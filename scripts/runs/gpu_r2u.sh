#!/bin/bash
# C2 shadow kernel, sustained (200 calls): single-CTA vs CTA-pair tiles, list lengths
mkdir -p gpurun_out
for cfg in "0 32" "1 32" "0 24" "0 48"; do set -- $cfg; echo -n "2CTA=$1 L=$2: "; AURA_GEMM_2CTA=$1 AURA_SHADOW_L=$2 python scripts/c2_shadow_one.py 200 2>&1 | tail -1; done | tee gpurun_out/r2u.log
