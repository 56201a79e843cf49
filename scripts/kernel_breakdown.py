"""Average device time per kernel of the C2 shadow batch call (CUPTI through torch.profiler; steady state, 100 calls).
usage: kernel_breakdown.py [iters] [rows]"""
import sys, json, torch
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
from aura_snn_rag_b200 import ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
rows = torch.randn(N, 768, device=dev, generator=g)
inv = ops.row_inv_norms(rows)
sh = ops.Bf16Shadow(rows)
q = rows[torch.randint(0, N, (1024,), device=dev, generator=g)] + 0.1 * torch.randn(1024, 768, device=dev, generator=g)
for _ in range(20):
    ops.batch_topk(rows, q, 10, inv, eps=1.0, shadow=sh)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(iters):
        ops.batch_topk(rows, q, 10, inv, eps=1.0, shadow=sh)
    torch.cuda.synchronize()
out = {}
for e in prof.key_averages():
    if e.device_time_total > 0:
        out[e.key[:60]] = round(e.device_time_total / iters, 2)
_, _, fl = ops.batch_topk(rows, q, 10, inv, eps=1.0, shadow=sh)
out["uncertain"] = int(fl.sum())
print(json.dumps(out))
