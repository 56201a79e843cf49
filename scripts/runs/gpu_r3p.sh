#!/bin/bash
# final sanity pass of HEAD: full GPU suite, smoke, C2 kernel breakdown
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02e_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02e_smoke.log
python scripts/kernel_breakdown.py 100 1000000 2>&1 | tail -1 | tee gpurun_out/r02e_c2_breakdown.json
