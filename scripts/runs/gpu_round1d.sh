#!/bin/bash
# full validation + contract bench + refreshed ncu evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_full.log | cut -c1-400
tail -3 gpurun_out/bench_err.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
python scripts/quick_tc_bench.py k7one > gpurun_out/plain_k7.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel -s 1 -c 1 -o gpurun_out/prof_k7 python scripts/quick_tc_bench.py k7one > gpurun_out/ncu_k7.log 2>&1
echo "ncu k7 rc=$?"
python scripts/quick_scan_bench.py one > gpurun_out/plain_scan.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 3 -c 1 -o gpurun_out/prof_scan python scripts/quick_scan_bench.py one > gpurun_out/ncu_scan.log 2>&1
echo "ncu scan rc=$?"
cat gpurun_out/plain_k7.log gpurun_out/plain_scan.log
