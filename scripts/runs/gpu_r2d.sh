#!/bin/bash
# round 2, call d: first run of the rows-as-M one-pass IVF kernel: parity tests, then C4 / C5-shard timings (new vs old)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tensorcore_gpu.py tests/test_round2_gpu.py tests/test_hippocampal_gpu.py tests/test_sharded_gpu.py -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/r2d_pytest.log
for K in 100 10; do
  timeout 200 python scripts/c5_shard_profile.py 12500000 16384 $K 64 3 2>&1 | tail -1 | tee gpurun_out/r2d_c5_new_k$K.log
  AURA_IVF_OLD=1 timeout 200 python scripts/c5_shard_profile.py 12500000 16384 $K 64 3 2>&1 | tail -1 | tee gpurun_out/r2d_c5_old_k$K.log
done
timeout 300 python bench.py --steps 3 --no-cpu-baseline --legs c4 2>gpurun_out/r2d_c4_err.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['c4_ivf']; print({k:(c[k]['ms_per_batch'] if isinstance(c[k],dict) and 'ms_per_batch' in c[k] else c[k]) for k in ['gather','list_major','recall_at_10_vs_exact','list_major_same_result','single_query_ms','batch_after_rebuild_ms']})" | tee gpurun_out/r2d_c4_new.log
tail -3 gpurun_out/r2d_c4_err.log
