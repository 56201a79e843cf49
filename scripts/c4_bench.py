"""BASELINE config 4 measurement (not the contract bench): IVF centroid index, M x 1024 fp32, 4096 centroids,
nprobe 32, batch 4096 queries, then 100k one-shot writes + incremental rebuild.  usage: c4_bench.py [M] [writes]"""
import json, sys, time, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.hippocampal import HippocampalFormation

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
D, C, P, B, K = 1024, 4096, 32, 4096, 10
dev = torch.device("cuda:0")

def timed(fn, iters=1, warm=0):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out

hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=M + W, feature_dim=D,
                          device="cuda:0", centroids_k=C, nprobe=P, track_ids=False)
hf.centroids_update_interval = 1 << 40
g = torch.Generator(device=dev).manual_seed(1234)
centres = torch.nn.functional.normalize(torch.randn(1024, D, device=dev, generator=g), dim=1)
def make_rows(n):
    which = torch.randint(0, 1024, (n,), device=dev, generator=g)
    return centres[which] + 0.05 * torch.randn(n, D, device=dev, generator=g)
t0 = time.time()
for r0 in range(0, M, 1 << 18):
    hf.create_episodic_memories(make_rows(min(1 << 18, M - r0)))
torch.cuda.synchronize()
res = {"M": M, "d": D, "C": C, "nprobe": P, "batch": B, "k": K, "fill_s": time.time() - t0}

seeds = torch.randperm(M, device=dev, generator=g)[:C]
ms, _ = timed(lambda: hf.rebuild_centroids(seed_rows=seeds))
res["rebuild_ms"] = ms
res["rebuild_tflops_assign_equiv"] = 2 * 2.0 * M * C * D / ms / 1e9
cnt = hf.centroid_counts
res["lists_min_max_mean"] = [float(cnt.min()), float(cnt.max()), float(cnt.mean())]

gq = torch.Generator(device=dev).manual_seed(4321)
pick = torch.randint(0, M, (B,), device=dev, generator=gq)
q = hf.memory_features[pick] + 0.1 * 0.05 * torch.randn(B, D, device=dev, generator=gq)
ms, (iv_idx, iv_sc) = timed(lambda: hf.retrieve_batch(q, K), iters=2, warm=1)
res["ivf_batch_ms"] = ms
res["ivf_qps"] = B / ms * 1e3
st = {}
hf._ensure_lists(); sc_, bi_ = hf._row_terms(None)
ops.ivf_search_batched(hf.memory_features, hf.memory_count, q, hf.centroids, P, hf._list_offsets, hf._list_rows, K, sc_, bi_,
                       eps=ops.TC_EPS_COS * 0.5, stats=st)
res["ivf_uncertified_queries"] = st["uncertain"]
probed = hf.centroid_counts[ops.ivf_coarse(q, hf.centroids, P).unique()].sum()
res["probed_list_bytes_GB"] = float(probed) * D * 4 / 1e9
res["ivf_list_bytes_per_s_TB"] = res["probed_list_bytes_GB"] / ms
ms, (ex_idx, ex_sc) = timed(lambda: hf.retrieve_batch(q, K, force_exact=True), iters=2, warm=1)
res["exact_batch_ms"] = ms
res["exact_qps"] = B / ms * 1e3
hits = (iv_idx.unsqueeze(2) == ex_idx.unsqueeze(1)).any(dim=2).float().sum(dim=1) / K
res["recall_at_10"] = float(hits.mean())
# list-major resident copy (2x bank memory): list tiles streamed by TMA instead of gathered
hf.list_major_copy = True
hf._ensure_lists()
hf._bank_by_list = torch.empty_like(hf.memory_features)
ms, _ = timed(lambda: ops.ivf_pack_lists(hf.memory_features, hf._list_rows, hf.memory_count, hf._bank_by_list), iters=2, warm=1)
res["pack_lists_ms"] = ms
ms, (lm_idx, lm_sc) = timed(lambda: hf.retrieve_batch(q, K), iters=2, warm=1)
res["ivf_batch_list_major_ms"] = ms
res["ivf_list_major_qps"] = B / ms * 1e3
res["ivf_list_major_bytes_per_s_TB"] = res["probed_list_bytes_GB"] / ms
res["list_major_same_result"] = bool(torch.equal(lm_idx, iv_idx) and torch.equal(lm_sc, iv_sc))
hf.list_major_copy = False
hf._bank_by_list = None
torch.cuda.empty_cache()
ms1, _ = timed(lambda: hf.retrieve_batch(q[:1], K), iters=20, warm=3)
res["ivf_single_query_ms"] = ms1
ms, _ = timed(lambda: ops.ivf_coarse(q, hf.centroids, P), iters=3, warm=1)
res["coarse_batch_ms"] = ms

new_rows = make_rows(W)
ms, _ = timed(lambda: hf.create_episodic_memories(new_rows))
res["online_writes"] = W
res["online_writes_ms"] = ms
res["online_writes_per_s"] = W / ms * 1e3
seeds = torch.randperm(hf.memory_count, device=dev, generator=g)[:C]
ms, _ = timed(lambda: hf.rebuild_centroids(seed_rows=seeds))
res["incremental_rebuild_ms"] = ms
ms, (iv_idx, _) = timed(lambda: hf.retrieve_batch(q, K), iters=1, warm=1)
res["ivf_batch_after_rebuild_ms"] = ms
print(json.dumps(res))
