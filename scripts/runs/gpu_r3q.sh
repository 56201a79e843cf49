#!/bin/bash
# two-depth second chance: full GPU suite, C2 kernel breakdown (1M rows and one rank's share at 8 GPUs), hand-back count
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02f_pytest_gpu.log
python scripts/kernel_breakdown.py 100 1000000 2>&1 | tail -1 | tee gpurun_out/r02f_c2_breakdown.json
python scripts/kernel_breakdown.py 100 125000 2>&1 | tail -1 | tee -a gpurun_out/r02f_c2_breakdown.json
