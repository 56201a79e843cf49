#!/bin/bash
# first GPU pass: parity tests, smoke, tuning sweep, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 300 python scripts/quick_scan_bench.py 2>&1 | tee gpurun_out/scan_sweep.log
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_short.log
