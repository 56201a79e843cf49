#!/bin/bash
# rows-as-M epilogue with pipelined 32-column TMEM loads: parity, C5 shard breakdown (resident-query mode off / on)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensorcore_gpu.py tests/test_round2_gpu.py tests/test_hippocampal_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r3c_pytest.log
for rb in 0 1; do echo -n "RB=$rb: "; AURA_IVF_RB=$rb timeout 600 python scripts/kernel_breakdown_c5.py 2>&1 | tail -1 | sed 's/.*scan_offsets[^]]*\], //'; done | tee gpurun_out/r3c.log
