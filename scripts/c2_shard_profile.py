"""One rank's share of the C2 bench at N GPUs (1M/N rows, batch 1024, k 10, bf16 shadow): step time eager / graphed,
for an ncu launch list.  usage: c2_shard_profile.py [rows] [iters]"""
import sys, json, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.sharded import ShardedBank
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
bank = torch.randn(n, 768, device=dev, generator=g)
inv = ops.row_inv_norms(bank)
sh = ops.Bf16Shadow(bank)
shard = ShardedBank(bank, 0, scale=inv, shadow=sh, peer_gather=True)      # world = 1: the exchange kernels still run
shard.world = 1
q = [bank[torch.randint(0, n, (1024,), device=dev, generator=g)] + 0.1 * torch.randn(1024, 768, device=dev, generator=g) for _ in range(4)]
def timed(fn):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
res = {"rows": n}
res["eager_ms"] = timed(lambda i: shard.search(q[i % 4], 10))
gs = [shard.graphed(1024, 10) for _ in range(2)]
res["graph_ms"] = timed(lambda i: shard.finalize(gs[i % 2].launch(q[i % 4])))
res["kernel_only_ms"] = timed(lambda i: ops.batch_topk(bank, q[i % 4], 10, inv, eps=1.0, shadow=sh))
print(json.dumps(res))
