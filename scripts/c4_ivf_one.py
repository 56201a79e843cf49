"""One C4-scale list-major IVF batch (for ncu).  usage: [LM=0|1|2] c4_ivf_one.py [M]   (LM=2: bf16 list-major shadow)"""
import sys, torch
sys.path.insert(0, ".")
from aura_snn_rag_b200.hippocampal import HippocampalFormation
M, D, C, P, B, K = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 1024, 4096, 32, 4096, 10
dev = torch.device("cuda:0")
hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=M, feature_dim=D, device="cuda:0",
                          centroids_k=C, nprobe=P, track_ids=False, list_major_copy={"0": False, "1": True, "2": "bf16"}[__import__("os").environ.get("LM", "0")])
hf.centroids_update_interval = 1 << 40
hf.ivf_strict = bool(int(__import__('os').environ.get('STRICT', '1')))
g = torch.Generator(device=dev).manual_seed(1234)
centres = torch.nn.functional.normalize(torch.randn(1024, D, device=dev, generator=g), dim=1)
for r0 in range(0, M, 1 << 18):
    n = min(1 << 18, M - r0)
    hf.create_episodic_memories(centres[torch.randint(0, 1024, (n,), device=dev, generator=g)] + 0.05 * torch.randn(n, D, device=dev, generator=g))
hf.rebuild_centroids(seed_rows=torch.randperm(M, device=dev, generator=g)[:C])
q = hf.memory_features[torch.randint(0, M, (B,), device=dev, generator=g)] + 0.005 * torch.randn(B, D, device=dev, generator=g)
for _ in range(3):
    hf.retrieve_batch(q, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); hf.retrieve_batch(q, K); e1.record(); torch.cuda.synchronize()
st = {}
from aura_snn_rag_b200 import ops
sc_, bi_ = hf._row_terms(None)
ops.ivf_search_batched(hf.memory_features, M, q, hf.centroids, P, hf._list_offsets, hf._list_rows, K, sc_, bi_,
                       eps=0.5 if hf._lm_bf16 else 0.5 * ops.TC_EPS_COS, stats=st, rows_by_list=hf._rows_by_list(), lm_relerr=hf._lm_relerr)
print(f"M={M} LM={__import__('os').environ.get('LM', '0')} ivf batch {e0.elapsed_time(e1):.2f} ms  handed back {st['handed_back']} of {B}")
