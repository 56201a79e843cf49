#!/bin/bash
# round 2, last validation pass (after the 256-query groups of the rows-as-M kernel, the reconverging TMEM load wrappers and
# the trace words): GPU tests, smoke, contract bench (all legs), ncu capture of the C5 fine-stage launches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02d_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02d_smoke.log
timeout 900 python bench.py 2>gpurun_out/r02d_bench_err.log > gpurun_out/r02d_bench_n1.json; echo "bench rc=$?"; tail -2 gpurun_out/r02d_bench_err.log
timeout 300 python scripts/c5_lm_one.py > gpurun_out/plain_c5_rows.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ivf_rows_kernel -s 9 -c 3 -f -o gpurun_out/r02d_c5_rows python scripts/c5_lm_one.py > gpurun_out/ncu_c5_rows.log 2>&1
echo "ncu c5_rows rc=$?"; tail -4 gpurun_out/plain_c5_rows.log
python scripts/kernel_breakdown.py 100 1000000 2>&1 | tail -1 | tee gpurun_out/r02d_c2_breakdown.json
