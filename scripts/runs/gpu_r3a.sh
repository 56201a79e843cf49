#!/bin/bash
# resident-query mode of the rows-as-M kernel: parity (both fine-stage formulations, list-major), C5 shard breakdown on/off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensorcore_gpu.py tests/test_round2_gpu.py tests/test_hippocampal_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r3a_pytest.log
for rb in 0 1; do echo -n "RB=$rb: "; AURA_IVF_RB=$rb timeout 600 python scripts/kernel_breakdown_c5.py 2>&1 | tail -1; done | tee gpurun_out/r3a.log
