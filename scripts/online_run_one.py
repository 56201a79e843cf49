"""The cooperative write-run kernel at the BASELINE config 4 centroid shape (4096 x 1024 fp32): n sequential one-shot
writes in one launch, for ncu / timing.  usage: online_run_one.py [n_writes] [repeats]"""
import sys, json, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
D, C = 1024, 4096
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
cent = torch.nn.functional.normalize(torch.randn(C, D, device=dev, generator=g), dim=1)
rows = cent[torch.randint(0, C, (n,), device=dev, generator=g)] + 0.05 * torch.randn(n, D, device=dev, generator=g)
counts = torch.full((C,), 100.0, device=dev)
cid = torch.empty(n, dtype=torch.int32, device=dev)
ops.online_assign(rows, 0, n, cent.clone(), C, counts.clone(), cid)
torch.cuda.synchronize()
ts = []
for _ in range(rep):
    c2, k2 = cent.clone(), counts.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.online_assign(rows, 0, n, c2, C, k2, cid); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(json.dumps({"writes": n, "ms": min(ts), "writes_per_s": n / min(ts) * 1e3, "us_per_write": min(ts) / n * 1e3}))
