"""batch_topk at one (dtype, B) for ncu.  usage: tc_small_one.py [f32|bf16] [B]"""
import sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
dt = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else torch.float32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
M, D, K = 1_000_000, 768, 10
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
rows = torch.randn(M, D, device=dev, generator=g).to(dt)
inv = ops.row_inv_norms(rows)
q = torch.randn(B, D, device=dev, generator=g)
for _ in range(4):
    ops.batch_topk(rows, q, K, inv, None)
torch.cuda.synchronize()
