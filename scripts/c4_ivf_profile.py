"""One BASELINE config 4 index (10M x 1024 fp32, 4096 lists, nprobe 32) and a few batch searches, for ncu launch lists /
--set full captures.  usage: c4_ivf_profile.py [K] [list_major 0|1] [iters] [M]"""
import json, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.hippocampal import HippocampalFormation
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
LM = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 2
M = int(sys.argv[4]) if len(sys.argv) > 4 else 10_000_000
D, C, P, B = 1024, 4096, 32, 4096
dev = torch.device("cuda:0")
hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=M, feature_dim=D,
                          device="cuda:0", centroids_k=C, nprobe=P, track_ids=False, list_major_copy=bool(LM))
hf.centroids_update_interval = 1 << 40
g = torch.Generator(device=dev).manual_seed(1234)
centres = torch.nn.functional.normalize(torch.randn(1024, D, device=dev, generator=g), dim=1)
for r0 in range(0, M, 1 << 18):
    n = min(1 << 18, M - r0)
    hf.create_episodic_memories(centres[torch.randint(0, 1024, (n,), device=dev, generator=g)] + 0.05 * torch.randn(n, D, device=dev, generator=g))
hf.rebuild_centroids(seed_rows=torch.randperm(M, device=dev, generator=g)[:C])
gq = torch.Generator(device=dev).manual_seed(4321)
pick = torch.randint(0, M, (B,), device=dev, generator=gq)
q = hf.memory_features[pick] + 0.005 * torch.randn(B, D, device=dev, generator=gq)
hf.retrieve_batch(q, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(ITERS):
    idx, sc = hf.retrieve_batch(q, K)
e1.record(); torch.cuda.synchronize()
st = {}
sc_, bi_ = hf._row_terms(None)
ops.ivf_search_batched(hf.memory_features, hf.memory_count, q, hf.centroids, P, hf._list_offsets, hf._list_rows, K, sc_, bi_,
                       eps=ops.TC_EPS_COS * 0.5, stats=st, strict=False, rows_by_list=hf._rows_by_list())
print(json.dumps({"k": K, "list_major": LM, "ms_per_batch": e0.elapsed_time(e1) / ITERS, "top1": float((idx[:, 0] == pick).float().mean()), "stats": st}))
