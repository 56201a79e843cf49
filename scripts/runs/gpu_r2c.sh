#!/bin/bash
# round 2, call c: contract bench with the new legs (N=1), each leg bounded separately while debugging
mkdir -p gpurun_out
for legs in "none" "c3" "c4" "c5"; do
  timeout 240 python bench.py --steps 5 --no-cpu-baseline --legs "$legs" 2>gpurun_out/r2c_err_$legs.log > gpurun_out/r2c_bench_$legs.json; echo "legs=$legs rc=$?"
  tail -4 gpurun_out/r2c_err_$legs.log
done
