#!/bin/bash
# 4 GPUs: is the slow first repetition of the C2 leg tied to the sampled start threshold?
mkdir -p gpurun_out
N=4
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
for smp in -1 0 -1; do
AURA_GEMM_SAMPLE=$smp timeout 300 $T bench.py --gpus $N --steps 20 --warmup 3 --legs '' 2>gpurun_out/r3g_err.log > gpurun_out/r3g.json; echo "SAMPLE=$smp rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r3g.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["ms_per_step_reps"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_reps"])
PY
done
