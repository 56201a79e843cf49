"""One C5 shard (12.5M x 768 bf16, 16384 lists, nprobe 64, k = 100, batch 4096), list-major copy: relaxed vs strict batch
time and how many queries the tensor-core pass hands back, with the worst-case and the measured certification bound.
usage: c5_strict_one.py [M]"""
import json, sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200 import ops
from aura_snn_rag_b200.hippocampal import HippocampalFormation

M = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
C, K, P, D, B = 16384, 100, 64, 768, 4096
dev = torch.device("cuda:0")
hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=M, feature_dim=D,
                          device="cuda:0", centroids_k=C, nprobe=P, bank_dtype=torch.bfloat16, track_ids=False, list_major_copy=True)
hf.centroids_update_interval = 1 << 40
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


def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


res = {"M": M}
for strict in (False, True):
    hf.ivf_strict = strict
    ms, (idx, sc) = timed(lambda: hf.retrieve_batch(q, K))
    res["strict" if strict else "relaxed"] = {"ms_per_batch": ms, "top1_is_source": float((idx[:, 0] == pick).float().mean())}
hf._ensure_lists(); sc_, bi_ = hf._row_terms(None)
for name, kw in (("worst_case_bound", dict(eps=ops.TC_EPS_COS * 0.5)), ("measured_bound", dict(eps=0.5, measured_eps=True))):
    st = {}
    ops.ivf_search_batched(hf.memory_features, M, q, hf.centroids, P, hf._list_offsets, hf._list_rows, K, sc_, bi_,
                           stats=st, strict=False, rows_by_list=hf._rows_by_list(), **kw)
    res[name] = {"handed_back": st["handed_back"], "path": st["path"]}
print(json.dumps(res))
