#!/bin/bash
for L in 48 32 24; do for R in 125000 1000000; do AURA_SHADOW_L=$L timeout 120 python scripts/c2_shard_profile.py $R 30 2>&1 | tail -1; done; done
AURA_SHADOW_L=24 timeout 200 python bench.py --steps 10 --no-cpu-baseline --legs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L24', round(d['value']), d['ms_per_step'], d['uncertified_queries_rerun'])"
AURA_SHADOW_L=32 timeout 200 python bench.py --steps 10 --no-cpu-baseline --legs none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('L32', round(d['value']), d['ms_per_step'], d['uncertified_queries_rerun'])"
