#!/bin/bash
# C5 shard rows-as-M kernel: sensitivity to pipeline depth (stages) - is the heavy section latency-bound?
mkdir -p gpurun_out
for st in 0 2; do echo -n "STAGES=$st: "; AURA_IVF_STAGES=$st timeout 600 python scripts/kernel_breakdown_c5.py 2>&1 | tail -1; done | tee gpurun_out/r2z.log
