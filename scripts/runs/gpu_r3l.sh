#!/bin/bash
# heavy lists in groups of 256 queries (M128 x N256 tiles): parity, C5 shard breakdown 128 vs 256
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensorcore_gpu.py tests/test_round2_gpu.py tests/test_hippocampal_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r3l_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3l_pytest.log
for gm in 128 256; do echo -n "GMAX=$gm: "; AURA_IVF_GMAX=$gm timeout 600 python scripts/kernel_breakdown_c5.py 2>&1 | tail -1 | sed 's/.*scan_offsets[^]]*\], //'; done | tee gpurun_out/r3l.log
