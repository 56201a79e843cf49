#!/bin/bash
mkdir -p gpurun_out
python scripts/quick_tc_bench.py k7one 2>&1 | tail -1 | tee gpurun_out/r3k_k7.log
timeout 300 python -m pytest tests/test_tensorcore_gpu.py -m gpu -x -q 2>&1 | tail -2
