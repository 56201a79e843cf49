#!/bin/bash
# bf16 list-major shadow: C4 batch time with the fp32 copy and the shadow, strict and relaxed, 24- and 32-entry lists
mkdir -p gpurun_out
for cfg in "1 1 0" "2 1 0" "2 0 0" "2 1 1" "2 0 1"; do set -- $cfg; LM=$1 STRICT=$2 AURA_IVF_SHADOW_SMALL=$3 python scripts/c4_ivf_one.py 10000000 2>&1 | tail -1 | sed "s/^/LM=$1 STRICT=$2 SMALL=$3: /"; done | tee gpurun_out/r2s_c4.log
