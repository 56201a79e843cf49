"""One C5 shard, list-major copy, relaxed: 4 batches (for ncu: the 10th ivf_rows_kernel launch is a steady-state heavy
section).  usage: c5_lm_one.py [M]"""
import sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200.hippocampal import HippocampalFormation
M = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
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
for _ in range(4):
    idx, sc = hf.retrieve_batch(q, K)
torch.cuda.synchronize()
# work-table statistics of the heavy section: lists, queries per list, rows
from aura_snn_rag_b200 import ops
probes = ops.ivf_coarse(q, hf.centroids, P)
nq = torch.bincount(probes.flatten(), minlength=hf.centroids.shape[0])
ln = (hf._list_offsets[1:] - hf._list_offsets[:-1]).long()
for lo, hi in ((1, 32), (33, 64), (65, 1 << 30)):
    m = (nq >= lo) & (nq <= hi) & (ln > 0)
    print(f"lists probed by {lo}..{hi} queries: {int(m.sum())} lists, {int(ln[m].sum())} rows, {int((ln[m] * nq[m]).sum()) / 1e9:.3f} G pairs, "
          f"max list {int(ln[m].max()) if m.any() else 0} rows, max queries {int(nq[m].max()) if m.any() else 0}")
print("top1 source", float((idx[:, 0] == pick).float().mean()))
