"""Single-query centroid-path latency breakdown at C4-like scale."""
import sys, time, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.hippocampal import HippocampalFormation
M, D, C, P, K = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 1024, 4096, 32, 10
dev = torch.device("cuda:0")
hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=M, feature_dim=D, device="cuda:0",
                          centroids_k=C, nprobe=P, track_ids=False)
hf.centroids_update_interval = 1 << 40
g = torch.Generator(device=dev).manual_seed(1234)
centres = torch.nn.functional.normalize(torch.randn(1024, D, device=dev, generator=g), dim=1)
for r0 in range(0, M, 1 << 18):
    n = min(1 << 18, M - r0)
    hf.create_episodic_memories(centres[torch.randint(0, 1024, (n,), device=dev, generator=g)] + 0.05 * torch.randn(n, D, device=dev, generator=g))
hf.rebuild_centroids(seed_rows=torch.randperm(M, device=dev, generator=g)[:C])
q = hf.memory_features[torch.randint(0, M, (64,), device=dev, generator=g)] + 0.005 * torch.randn(64, D, device=dev, generator=g)
sc, bi = hf._row_terms(None)
def ev_time(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_coarse = ev_time(lambda i=0: ops.ivf_coarse(q[i % 64:i % 64 + 1], hf.centroids, P))
t_search = ev_time(lambda i=0: ops.ivf_search(hf.memory_features, M, q[i % 64:i % 64 + 1], hf.centroids, P, hf._list_offsets, hf._list_rows, K, sc, bi))
probes = ops.ivf_coarse(q[:1], hf.centroids, P)
cand = int(hf.centroid_counts[probes[0]].sum())
t0 = time.perf_counter()
for i in range(100): hf.retrieve_similar_memories(q[i % 64], k=K)
t_api = (time.perf_counter() - t0) / 100 * 1e6
print(f"M={M}: coarse {t_coarse:.1f} us, ivf_search (coarse+fine) {t_search:.1f} us, API {t_api:.1f} us; candidates of query 0: {cand} rows = {cand*D*4/1e6:.1f} MB "
      f"-> fine stage {cand*D*4/max(t_search-t_coarse,1)/1e3:.0f} GB/s")
