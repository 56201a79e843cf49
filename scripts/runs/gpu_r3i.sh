#!/bin/bash
# second chance in the queries-as-M IVF finish: parity, C4 with 24- / 32-entry shadow lists
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensorcore_gpu.py tests/test_round2_gpu.py tests/test_hippocampal_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r3i_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3i_pytest.log
for cfg in "1 1 0" "2 1 0" "2 1 1" "2 0 1"; do set -- $cfg; LM=$1 STRICT=$2 AURA_IVF_SHADOW_SMALL=$3 python scripts/c4_ivf_one.py 10000000 2>&1 | tail -1 | sed "s/^/LM=$1 STRICT=$2 SMALL=$3: /"; done | tee gpurun_out/r3i_c4.log
