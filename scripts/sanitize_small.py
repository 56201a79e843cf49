"""Tiny invocation of every kernel family (for compute-sanitizer memcheck)."""
import sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.hippocampal import HippocampalFormation
torch.manual_seed(0)
dev = "cuda:0"
hf = HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=3000, feature_dim=100,
                          centroids_k=16, centroid_rows=20, nprobe=4)
hf.centroids_update_interval = 1 << 30
rows = torch.randn(2500, 100)
hf.create_episodic_memories(rows[:2400], [f"m{i}" for i in range(2400)])
hf.rebuild_centroids()
for i in range(2400, 2500):
    hf.create_episodic_memory(f"m{i}", "e", rows[i])
q = rows[:70] + 0.1 * torch.randn(70, 100)
print(hf.retrieve_similar_memories(q[0], k=5)[:2])
print(hf.retrieve_similar_memories(q[1], location=torch.tensor([0.5, 0.5]), k=5)[:1])
i1, s1 = hf.retrieve_batch(q, 10)                 # list-major IVF (B >= 64)
i2, s2 = hf.retrieve_batch(q[:9], 10)             # per-query IVF
i3, s3 = hf.retrieve_batch(q, 10, force_exact=True)   # K6
i4, s4 = hf.exact_topk(q[:3], 7)                  # scan
nbr, sim = hf.build_cognitive_map(8)              # K7
hf.decay_memories(0.1)
bf = HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=3000, feature_dim=128,
                          bank_dtype=torch.bfloat16, use_centroid_index=False, track_ids=False)
bf.create_episodic_memories(torch.randn(2048, 128))
bf.exact_topk(torch.randn(16, 128), 10); bf.exact_topk(torch.randn(1, 128), 10); bf.build_cognitive_map(4)
a = torch.empty(2400, dtype=torch.int32, device=dev)
import os
os.environ["AURA_ASSIGN_TC"] = "1"
ops.kmeans_assign(hf.memory_features, 2400, hf.centroids, 16, a, inv_norm=hf._inv_norm)
torch.cuda.synchronize()
print("sanitize run ok", i1.shape, i3.shape, nbr.shape)
