#!/bin/bash
# round 2, call n (8 GPUs): the contract bench as the driver launches it at N=8, plus N=4
mkdir -p gpurun_out
for N in 8 4; do
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
timeout 400 $T bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/r2n_err_n$N.log > gpurun_out/r2n_n$N.json; echo "N=$N rc=$?"
grep -E "bench|Error|error" gpurun_out/r2n_err_n$N.log | tail -5
python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_n$N.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["ms_per_step_reps"], "e2e", round(d["e2e"]["value"]), d["exchange"], d["cuda_graph_step"], d["top1_hit_rate"])
c=d.get("c5_sharded"); print({k:c[k] for k in ["build_s","relaxed","strict","recall_at_10_vs_exact","recall_at_100_vs_exact","list_major_copy"]}, c["roofline"]["frac"])
PY
done
