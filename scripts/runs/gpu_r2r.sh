#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tensorcore_gpu.py tests/test_round2_gpu.py tests/test_hippocampal_gpu.py tests/test_sharded_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 200 python scripts/c4_ivf_profile.py 10 1 3 2>&1 | tail -1
timeout 200 python scripts/c4_ivf_profile.py 10 0 3 2>&1 | tail -1
timeout 200 python scripts/c5_shard_profile.py 12500000 16384 10 64 3 2>&1 | tail -1
