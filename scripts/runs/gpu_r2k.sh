#!/bin/bash
# round 2, call k (2 GPUs): peer-memory exchange + graphed sharded step vs the NCCL path
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_sharded_gpu.py tests/test_round2_gpu.py -q -m gpu -x -k "peer_gather or non_current" 2>&1 | tail -3
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 2 --steps 20 --legs none 2>gpurun_out/r2k_err_peer.log > gpurun_out/r2k_n2_peer.json; echo "peer rc=$?"; tail -4 gpurun_out/r2k_err_peer.log
timeout 300 $T bench.py --gpus 2 --steps 20 --legs none --no-peer-gather 2>gpurun_out/r2k_err_nccl.log > gpurun_out/r2k_n2_nccl.json; echo "nccl rc=$?"; tail -2 gpurun_out/r2k_err_nccl.log
python - <<'PY'
import json
for f in ["peer","nccl"]:
    try:
        d=json.load(open(f"gpurun_out/r2k_n2_{f}.json"))
        print(f, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["exchange"], d["cuda_graph_step"], d["gpu_launches"], d["top1_hit_rate"])
    except Exception as e: print(f, "failed", e)
PY
