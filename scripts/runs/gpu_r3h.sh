#!/bin/bash
# second-chance certificate: full GPU suite, C2 hand-backs over 50 distinct batches, C5 shard strict vs relaxed
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3h_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r3h_pytest.log
python - <<'PY' 2>&1 | tail -3
import torch, sys
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1234)
rows = torch.randn(1_000_000, 768, device=dev, generator=g)
inv = ops.row_inv_norms(rows)
sh = ops.Bf16Shadow(rows)
tot = 0
for it in range(50):
    q = rows[torch.randint(0, 1_000_000, (1024,), device=dev, generator=g)] + 0.1 * torch.randn(1024, 768, device=dev, generator=g)
    _, _, fl = ops.batch_topk(rows, q, 10, inv, eps=1.0, shadow=sh)
    tot += int(fl.sum())
print("C2 shadow L=32: uncertified of 51200:", tot)
PY
python scripts/c5_strict_one.py 2>&1 | tail -1 | tee gpurun_out/r3h_c5.json
