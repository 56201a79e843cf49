"""Exact search of small query blocks: streaming scan (fp32 FMA, QB <= 8 per pass) vs tensor-core paths."""
import os, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
M, D, K = 1_000_000, 768, 10
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
one = len(sys.argv) > 1 and sys.argv[1] == "one"
def t(fn, n=6):
    ts = []
    for i in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts[1:])
for dt in (torch.float32,) if one else (torch.float32, torch.bfloat16):
    rows = torch.randn(M, D, device=dev, generator=g).to(dt)
    inv = ops.row_inv_norms(rows)
    for B in (16,) if one else (1, 2, 3, 4, 8, 16, 32, 64, 128, 256):
        q = torch.randn(B, D, device=dev, generator=g)
        ts = t(lambda: ops.scan_topk(rows, q, K, inv, None))
        tb = t(lambda: ops.batch_topk(rows, q, K, inv, None))
        print(f"{str(dt)[6:]} B={B}: scan {ts:.0f} us, tensor-core {tb:.0f} us", flush=True)
