#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_ivf_scale_gpu.py 2>&1 | tail -6 | tee gpurun_out/r2p_pytest.log
for R in 125000 1000000; do timeout 120 python scripts/c2_shard_profile.py $R 30 2>&1 | tail -1; AURA_GEMM_GTHR=0 timeout 120 python scripts/c2_shard_profile.py $R 30 2>&1 | tail -1; done
timeout 200 python scripts/sb_vs_scan.py 2>&1 | tail -24
timeout 300 python bench.py --steps 10 --no-cpu-baseline --legs c3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['c3_allpairs']['ms'], d['c3_allpairs']['roofline']['frac'])"
