#!/bin/bash
# measured bound for bf16 banks + 128-key shortlists at k = 100: parity tests, one C5 shard strict vs relaxed
mkdir -p gpurun_out
python -m pytest tests/test_tensorcore_gpu.py tests/test_hippocampal_gpu.py tests/test_round2_gpu.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2t_pytest.log
python scripts/c5_strict_one.py 2>&1 | tail -1 | tee gpurun_out/r2t_c5.json
