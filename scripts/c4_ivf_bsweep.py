"""IVF list-major batch at C4 scale for several batch sizes: probed list bytes vs time (gather / list-major copy)."""
import os, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.hippocampal import HippocampalFormation
M, D, C, P, K = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 1024, 4096, 32, 10
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
ln = (hf._list_offsets[1:] - hf._list_offsets[:-1]).float()
for B in (256, 512, 1024, 2048, 4096, 8192):
    q = hf.memory_features[torch.randint(0, M, (B,), device=dev, generator=g)] + 0.005 * torch.randn(B, D, device=dev, generator=g)
    pr = ops.ivf_coarse(q, hf.centroids, P)
    nq = torch.bincount(pr.flatten().long(), minlength=ln.numel()).float()
    once = (ln * (nq > 0)).sum().item() * D * 4 / 1e9
    tiles = (ln * torch.ceil(nq / 128)).sum().item() * D * 4 / 1e9
    out = []
    for lm in (0, 1):
        hf.list_major_copy = bool(lm)
        hf.retrieve_batch(q, K)
        ts = []
        for i in range(4):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); hf.retrieve_batch(q, K); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out.append(min(ts))
    print(f"B={B}: probed once {once:.1f} GB, x query tiles {tiles:.1f} GB; gather {out[0]:.2f} ms ({once/out[0]:.2f} TB/s), list-major {out[1]:.2f} ms ({once/out[1]:.2f} TB/s)", flush=True)
