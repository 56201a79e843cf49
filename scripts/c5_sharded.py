"""BASELINE config 5 measurement: bf16 bank row-sharded over the ranks of one box, IVF nprobe 64, k = 100,
NCCL all-gather top-k merge.  Launch: torchrun --nproc-per-node N scripts/c5_sharded.py [M_total] [C]
(single process: python scripts/c5_sharded.py M_total)."""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from aura_snn_rag_b200.hippocampal import HippocampalFormation
from aura_snn_rag_b200.sharded import ShardedIndex, shard_range

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
D, P, K, B = 768, 64, 100, 4096
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lrank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

lo, hi = shard_range(M, rank, world)
m = hi - lo
hf = HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=m, feature_dim=D,
                          device=f"cuda:{lrank}", centroids_k=C, nprobe=P, bank_dtype=torch.bfloat16, track_ids=False)
hf.centroids_update_interval = 1 << 40
gc = torch.Generator(device=dev).manual_seed(99)                     # same cluster centres on every rank
centres = torch.nn.functional.normalize(torch.randn(8192, D, device=dev, generator=gc), dim=1)
g = torch.Generator(device=dev).manual_seed(1234 + rank)
t0 = time.time()
for r0 in range(0, m, 1 << 18):
    n = min(1 << 18, m - r0)
    hf.create_episodic_memories(centres[torch.randint(0, 8192, (n,), device=dev, generator=g)] +
                                0.05 * torch.randn(n, D, device=dev, generator=g))
barrier()
res = {"M_total": M, "ranks": world, "rows_per_rank": m, "d": D, "C": C, "nprobe": P, "k": K, "batch": B,
       "fill_s": time.time() - t0}

idx = ShardedIndex(hf, lo, M)
seeds = torch.randperm(M, device=dev, generator=torch.Generator(device=dev).manual_seed(7))[:C]   # same on all ranks
barrier(); t0 = time.time()
idx.rebuild_centroids(seeds)
barrier(); res["rebuild_s"] = time.time() - t0

# queries = stored row (fetched from its owner) + noise, identical on all ranks
gq = torch.Generator(device=dev).manual_seed(4321)
pick = torch.randint(0, M, (B,), device=dev, generator=gq)
q = torch.zeros(B, D, device=dev)
own = (pick >= lo) & (pick < hi)
q[own] = hf.memory_features[(pick[own] - lo)].float()
if world > 1:
    dist.all_reduce(q)
q += 0.005 * torch.randn(B, D, device=dev, generator=gq)

def timed(fn, iters):
    fn(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), out

ms, (ii, ss) = timed(lambda: idx.search(q, K), 2)
res["ivf_batch_ms"] = ms; res["ivf_qps"] = B / ms * 1e3
hf.ivf_strict = False
ms2, (ii2, ss2) = timed(lambda: idx.search(q, K), 3)
res["ivf_relaxed_batch_ms"] = ms2; res["ivf_relaxed_qps"] = B / ms2 * 1e3
res["relaxed_vs_strict_overlap_at_100"] = float((ii2.unsqueeze(2) == ii.unsqueeze(1)).any(dim=2).float().mean())
nq = 256
ms_e, (ie, se) = timed(lambda: idx.search(q[:nq], K, exact=True), 1)
res["exact_batch256_ms"] = ms_e
hits = (ii[:nq].unsqueeze(2) == ie.unsqueeze(1)).any(dim=2).float().sum(dim=1) / K
res["recall_at_100"] = float(hits.mean())
hits10 = (ii[:nq, :10].unsqueeze(2) == ie[:, :10].unsqueeze(1)).any(dim=2).float().sum(dim=1) / 10
res["recall_at_10"] = float(hits10.mean())
res["top1_is_source_row"] = float((ii[:, 0] == pick).float().mean())
res["list_bytes_per_rank_GB"] = m * D * 2 / 1e9
from aura_snn_rag_b200 import ops
st = {}
sc_, bi_ = hf._row_terms(None)
ops.ivf_search_batched(hf.memory_features, hf.memory_count, q, hf.centroids, P, hf._list_offsets, hf._list_rows, K, sc_, bi_,
                       eps=ops.TC_EPS_COS * 0.5, stats=st)
res["uncertified_on_rank0"] = st["uncertain"]; res["items"] = st.get("items"); res["items_cap"] = st.get("items_cap")
cnt = (hf._list_offsets[1:] - hf._list_offsets[:-1]).float()
res["local_list_len_min_mean_max"] = [float(cnt.min()), float(cnt.mean()), float(cnt.max())]
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
