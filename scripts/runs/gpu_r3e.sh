#!/bin/bash
# seeded start bound of the list-major fine stages: parity, C5 shard breakdown with / without, C4 batch with / without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensorcore_gpu.py tests/test_round2_gpu.py tests/test_hippocampal_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r3e_pytest.log
for sd in 0 1; do echo -n "C5 SEED=$sd: "; AURA_IVF_RB=0 AURA_IVF_SEED=$sd timeout 600 python scripts/kernel_breakdown_c5.py 2>&1 | tail -1 | sed 's/.*scan_offsets[^]]*\], //'; done | tee gpurun_out/r3e.log
echo -n "C5 SEED=1 RB=1: "; AURA_IVF_RB=1 timeout 600 python scripts/kernel_breakdown_c5.py 2>&1 | tail -1 | sed 's/.*scan_offsets[^]]*\], //' | tee -a gpurun_out/r3e.log
for sd in 0 1; do for lm in 1 2; do echo -n "C4 SEED=$sd "; AURA_IVF_SEED=$sd LM=$lm STRICT=1 timeout 600 python scripts/c4_ivf_one.py 10000000 2>&1 | tail -1; done; done | tee -a gpurun_out/r3e.log
