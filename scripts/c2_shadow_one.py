"""The C2 headline kernel alone: batch 1024 against 1M x 768 fp32 through the bf16 shadow (k = 10), for ncu.
usage: c2_shadow_one.py [iters]"""
import sys, json, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
rows = torch.randn(1_000_000, 768, device=dev, generator=g)
inv = ops.row_inv_norms(rows)
sh = ops.Bf16Shadow(rows)
q = rows[torch.randint(0, 1_000_000, (1024,), device=dev, generator=g)] + 0.1 * torch.randn(1024, 768, device=dev, generator=g)
for _ in range(2):
    ops.batch_topk(rows, q, 10, inv, eps=1.0, shadow=sh)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    _, _, fl = ops.batch_topk(rows, q, 10, inv, eps=1.0, shadow=sh)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"ms_per_call": e0.elapsed_time(e1) / iters, "uncertain": int(fl.sum())}))
