"""Average device time per kernel of one relaxed C5-shard batch (12.5M x 768 bf16, 16384 lists, nprobe 64, k = 100,
batch 4096, list-major copy), CUPTI through torch.profiler.  usage: kernel_breakdown_c5.py [M] [iters]"""
import sys, json, torch
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
from aura_snn_rag_b200.hippocampal import HippocampalFormation
M = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
C, K, P, D, B = 16384, 100, 64, 768, 4096
dev = torch.device("cuda:0")
hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=M, feature_dim=D,
                          device="cuda:0", centroids_k=C, nprobe=P, bank_dtype=torch.bfloat16, track_ids=False, list_major_copy=True)
hf.centroids_update_interval = 1 << 40
hf.ivf_strict = False
gc = torch.Generator(device=dev).manual_seed(99)
centres = torch.nn.functional.normalize(torch.randn(8192, D, device=dev, generator=gc), dim=1)
g = torch.Generator(device=dev).manual_seed(1234)
for r0 in range(0, M, 1 << 18):
    n = min(1 << 18, M - r0)
    hf.create_episodic_memories(centres[torch.randint(0, 8192, (n,), device=dev, generator=g)] +
                                0.05 * torch.randn(n, D, device=dev, generator=g))
hf.rebuild_centroids(seed_rows=torch.randperm(M, device=dev, generator=torch.Generator(device=dev).manual_seed(7))[:C])
gq = torch.Generator(device=dev).manual_seed(4321)
pick = torch.randint(0, M, (B,), device=dev, generator=gq)
q = hf.memory_features[pick].float() + 0.005 * torch.randn(B, D, device=dev, generator=gq)
for _ in range(3):
    hf.retrieve_batch(q, K)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(iters):
        hf.retrieve_batch(q, K)
    torch.cuda.synchronize()
out = {}
for e in prof.key_averages():
    if e.device_time_total / iters > 20:
        out[e.key[:48]] = [round(e.device_time_total / iters, 1), e.count // iters]
print(json.dumps(out))
