#!/bin/bash
mkdir -p gpurun_out
for h in 3 259 771; do echo -n "L2HINT=$h: "; AURA_IVF_RB=0 AURA_IVF_L2HINT=$h timeout 600 python scripts/kernel_breakdown_c5.py 2>&1 | tail -1 | sed 's/.*scan_offsets[^]]*\], //'; done | tee gpurun_out/r3d.log
