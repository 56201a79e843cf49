"""IVF list-major tuning sweep at C4 scale (experiments)."""
import os, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.hippocampal import HippocampalFormation
M, D, C, P, B, K = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 1024, 4096, 32, 4096, 10
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
pick = torch.randint(0, M, (B,), device=dev, generator=g)
q = hf.memory_features[pick] + 0.005 * torch.randn(B, D, device=dev, generator=g)
ref = None
for i in range(3):   # box sanity: exact search of 1024 queries (known: ~33 ms at 10M x 1024)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); hf.exact_topk(q[:1024], K); e1.record(); torch.cuda.synchronize()
    print("exact 1024 queries ms", round(e0.elapsed_time(e1), 2), flush=True)
# list statistics: how many query tiles re-read each list
probes = ops.ivf_coarse(q, hf.centroids, P) if hasattr(ops, "ivf_coarse") else None
if probes is not None:
    pr = probes[0] if isinstance(probes, tuple) else probes
    nq = torch.bincount(pr.flatten().long(), minlength=C).float()
    ln = (hf._list_offsets[1:] - hf._list_offsets[:-1]).float() if hasattr(hf, "_list_offsets") else None
    if ln is not None:
        once = (ln * (nq > 0)).sum().item(); t128 = (ln * torch.ceil(nq / 128)).sum().item()
        print(f"list rows probed once {once:.3e}, x query tiles {t128:.3e} ({t128/once:.2f}x), lists with >128 queries {(nq > 128).sum().item()}, max nq {nq.max().item():.0f}, max len {ln.max().item():.0f}", flush=True)
for env in [dict(LM=c) for c in (0,1,0,1,0,1)]:
    os.environ.update({k: str(v) for k, v in env.items()})
    hf.list_major_copy = bool(env.get('LM', 0))
    hf.retrieve_batch(q, K)
    ts = []
    for i in range(6):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); hf.retrieve_batch(q, K); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    r = hf.retrieve_batch(q, K)
    if ref is None: ref = r
    same = torch.equal(r[0], ref[0])
    print(env, " ".join(f"{t:.1f}" for t in ts), "same_as_first", same, flush=True)
