#!/bin/bash
# parity + contract bench + ncu evidence (launch list of the bench, full sets of the two top kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_full.log | cut -c1-1500
tail -3 gpurun_out/bench_err.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
python scripts/quick_tc_bench.py k6one > gpurun_out/plain_k6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel -s 2 -c 1 -o gpurun_out/prof_k6 python scripts/quick_tc_bench.py k6one > gpurun_out/ncu_k6.log 2>&1
echo "ncu k6 rc=$?"
python scripts/quick_scan_bench.py one > gpurun_out/plain_scan.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 3 -c 1 -o gpurun_out/prof_scan python scripts/quick_scan_bench.py one > gpurun_out/ncu_scan.log 2>&1
echo "ncu scan rc=$?"
ls -la gpurun_out
