#!/bin/bash
# full validation + contract bench + refreshed evidence (run under gpurun, one GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_full.log | cut -c1-600
tail -3 gpurun_out/bench_err.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
timeout 300 python scripts/sb_vs_scan.py 2>&1 | tee gpurun_out/small_blocks.log | tail -20
timeout 600 python scripts/c4_bench.py 10000000 100000 2>&1 | tail -1 | tee gpurun_out/c4.json | cut -c1-1500
timeout 300 python scripts/c4_ivf_one.py 10000000 > gpurun_out/plain_ivf.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ivf_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_ivf python scripts/c4_ivf_one.py 10000000 > gpurun_out/ncu_ivf.log 2>&1
echo "ncu ivf rc=$?"
cat gpurun_out/plain_ivf.log
