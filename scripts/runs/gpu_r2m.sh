#!/bin/bash
# round 2, call m (2 GPUs): the contract bench exactly as the driver launches it at N=2 (default legs: c5_sharded 100M rows)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/r2m_err.log > gpurun_out/r2m_n2.json; echo "rc=$?"
grep -E "bench|Error|error" gpurun_out/r2m_err.log | tail -12
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2m_n2.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["exchange"], d["cuda_graph_step"])
print(json.dumps(d.get("c5_sharded"), indent=0)[:1800])
PY
