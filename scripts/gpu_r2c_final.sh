#!/bin/bash
# round 2, final validation pass: GPU tests, smoke, contract bench (all legs), reference arm, launch list of the bench
# command, ncu --set full captures of the kernel variants the bench now runs (each after the same command exited 0 without ncu)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02c_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02c_smoke.log
timeout 900 python bench.py 2>gpurun_out/r02c_bench_err.log > gpurun_out/r02c_bench_n1.json; echo "bench rc=$?"; tail -2 gpurun_out/r02c_bench_err.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02c_bench_ref_n1.json; cut -c1-300 gpurun_out/r02c_bench_ref_n1.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --legs none"
timeout 200 $B > gpurun_out/plain_bench.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'aura|gemm_topk|scan_topk|normalize|merge|pack' --csv --log-file gpurun_out/r02c_launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 300 "$@" > gpurun_out/plain_$name.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/r02c_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"; tail -1 gpurun_out/plain_$name.log | cut -c1-300
}
cap c2_shadow gemm_topk_kernel 2 1 python scripts/c2_shadow_one.py 2
cap k7 gemm_topk_kernel 1 1 python scripts/quick_tc_bench.py k7one
LM=1 cap c4_lm ivf_gemm_kernel 3 1 python scripts/c4_ivf_one.py 10000000
LM=2 cap c4_lm_bf16 ivf_gemm_kernel 3 1 python scripts/c4_ivf_one.py 10000000
cap c5_rows ivf_rows_kernel 9 3 python scripts/c5_lm_one.py
