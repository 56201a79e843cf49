"""End-to-end latency of the reference-facing calls (python -> C ABI -> python), BASELINE config 1 and 2 scale."""
import sys, time, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200.hippocampal import HippocampalFormation
dev = "cuda:0"
for n, d, index in ((10_000, 768, True), (10_000, 768, False), (1_000_000, 768, False)):
    hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=n, feature_dim=d, device=dev,
                              use_centroid_index=index)
    rows = torch.randn(n, d, device=dev)
    t0 = time.perf_counter()
    if n <= 10_000:
        for i in range(n):
            hf.create_episodic_memory(f"m{i}", "e", rows[i])
        torch.cuda.synchronize()
        t_ins = time.perf_counter() - t0
        print(f"n={n} index={index}: {n} create_episodic_memory calls {t_ins:.2f} s -> {n/t_ins:.0f} inserts/s (incl. {n//512} rebuilds)")
    else:
        hf.create_episodic_memories(rows, None)
        hf.track_ids = False
    q = rows[:200] + 0.1 * torch.randn(200, d, device=dev)
    qc = q.cpu()
    for name, qq in (("device query", q), ("host query", qc)):
        for i in range(10):
            hf.retrieve_similar_memories(qq[i], k=10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(200):
            r = hf.retrieve_similar_memories(qq[i], k=10)
        dt = (time.perf_counter() - t0) / 200
        print(f"n={n} index={index} centroid_path={hf._centroid_path()} {name}: retrieve_similar_memories {dt*1e3:.3f} ms/query -> {1/dt:.0f} qps")
    t0 = time.perf_counter()
    hf.rebuild_centroids() if index else None
    torch.cuda.synchronize()
    print(f"   rebuild_centroids: {(time.perf_counter()-t0)*1e3:.2f} ms")
