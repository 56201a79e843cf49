"""One C5 shard (BASELINE config 5 per-GPU slice at 8 GPUs: 12.5M x 768 bf16, 16384 lists, nprobe 64, k = 100, batch 4096):
one relaxed batch search, for an ncu launch list / event timings.  usage: c5_shard_profile.py [M] [C] [K] [P] [iters]"""
import json, sys, time, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.hippocampal import HippocampalFormation

M = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
P = int(sys.argv[4]) if len(sys.argv) > 4 else 64
ITERS = int(sys.argv[5]) if len(sys.argv) > 5 else 3
D, B = 768, 4096
dev = torch.device("cuda:0")
hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=M, feature_dim=D,
                          device="cuda:0", centroids_k=C, nprobe=P, bank_dtype=torch.bfloat16, track_ids=False)
hf.centroids_update_interval = 1 << 40
gc = torch.Generator(device=dev).manual_seed(99)
centres = torch.nn.functional.normalize(torch.randn(8192, D, device=dev, generator=gc), dim=1)
g = torch.Generator(device=dev).manual_seed(1234)
for r0 in range(0, M, 1 << 18):
    n = min(1 << 18, M - r0)
    hf.create_episodic_memories(centres[torch.randint(0, 8192, (n,), device=dev, generator=g)] +
                                0.05 * torch.randn(n, D, device=dev, generator=g))
seeds = torch.randperm(M, device=dev, generator=torch.Generator(device=dev).manual_seed(7))[:C]
hf.rebuild_centroids(seed_rows=seeds)
gq = torch.Generator(device=dev).manual_seed(4321)
pick = torch.randint(0, M, (B,), device=dev, generator=gq)
q = hf.memory_features[pick].float() + 0.005 * torch.randn(B, D, device=dev, generator=gq)
hf.ivf_strict = False
res = {"M": M, "C": C, "k": K, "nprobe": P, "batch": B}
for lm in (False, True):
    hf.list_major_copy = lm
    hf.retrieve_batch(q, K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        idx, sc = hf.retrieve_batch(q, K)
    e1.record(); torch.cuda.synchronize()
    res["list_major" if lm else "gather"] = {"ms_per_batch": e0.elapsed_time(e1) / ITERS,
                                             "top1_is_source": float((idx[:, 0] == pick).float().mean())}
st = {}
hf._ensure_lists(); sc_, bi_ = hf._row_terms(None)
ops.ivf_search_batched(hf.memory_features, hf.memory_count, q, hf.centroids, P, hf._list_offsets, hf._list_rows, K, sc_, bi_,
                       eps=ops.TC_EPS_COS * 0.5, stats=st, strict=False)
res["stats"] = st
res["list_bytes_GB"] = M * D * 2 / 1e9
cnt = (hf._list_offsets[1:] - hf._list_offsets[:-1]).float()
res["list_len_min_mean_max"] = [float(cnt.min()), float(cnt.mean()), float(cnt.max())]
print(json.dumps(res))
