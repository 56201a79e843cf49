#!/bin/bash
# final validation: GPU tests, smoke, contract bench, ncu launch list of the bench command (one GPU, tight timeouts)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 300 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_full.log | cut -c1-400
tail -3 gpurun_out/bench_err.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 120 $B > gpurun_out/plain.log 2>&1 &&
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
timeout 120 python scripts/sb_vs_scan.py 2>&1 | tee gpurun_out/small_blocks.log | tail -20
