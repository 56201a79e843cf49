"""scan kernel: fp32 / bf16, d=768 / 1024, single query and small batches (after consumer restructuring)."""
import os, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, e in evs:
        a.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(e) for a, e in evs)
    return ts[len(ts) // 2]
for n, d in ((1_000_000, 768), (1_000_000, 1024)):
    for dt in (torch.float32, torch.bfloat16):
        rows = torch.randn(n, d, device=dev).to(dt)
        inv = ops.row_inv_norms(rows)
        for b, k, env in ((1, 10, {}), (1, 10, {"AURA_SCAN_STAGES": 4}), (1, 10, {"AURA_SCAN_RU2": 1}), (2, 10, {}), (4, 10, {}), (1, 100, {})):
            for kk in ("AURA_SCAN_STAGES", "AURA_SCAN_RU2"): os.environ.pop(kk, None)
            os.environ.update({a: str(v) for a, v in env.items()})
            q = torch.randn(b, d, device=dev)
            med = timeit(lambda: ops.scan_topk(rows, q, k, scale=inv))
            byt = rows.numel() * rows.element_size()
            print(f"n={n} d={d} {str(dt)[6:]:8s} B={b} k={k:3d} env={env}: {med*1e3:7.1f} us {byt/med/1e6:6.0f} GB/s", flush=True)
        del rows
