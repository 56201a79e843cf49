"""Row-sharded memory bank: one process per GPU, local exact / IVF top-k, NCCL all-gather, k-way merge.

The reference is single-process; scoring of a bank row is independent of all
other rows (src/core/hippocampal.py:279,301-303), so the bank shards by rows.  Rank r owns the
contiguous global rows [row_base, row_base + n_local).  A query batch (replicated on every rank) is
answered by: local top-k with GLOBAL row ids -> all_gather of (score fp32, id int64) blocks ->
`aura_topk_merge` on every rank.  The merge of exact local top-k lists is the exact global top-k, and
the tie rule (lower global row first) is the same in the local scan and the merge, so the sharded
result equals the single-GPU result bit for bit.

`local_search` / `merge` are injectable so the rendezvous / layout logic can be exercised with the
gloo backend on CPU (tests/test_sharded_gloo.py); the defaults are the CUDA kernels and nothing else.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_rows over `world` ranks, remainder spread over the first ranks."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedBank:
    def __init__(self, local_rows: torch.Tensor, row_base: int, group=None,
                 scale: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 stats: Optional[dict] = None):
        self.rows = local_rows
        self.row_base = int(row_base)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.stats = stats if stats is not None else {}
        self._fixup = None
        self._packed = local_search is None and merge is None
        self._ops = None
        if local_search is None or merge is None:
            from . import ops      # CUDA only; raises without the library
            self._ops = ops
            if scale is None:
                scale = ops.row_inv_norms(local_rows)
            if local_search is None:
                local_search, self._fixup = self._default_search(ops)
            merge = merge or ops.topk_merge
        self.scale, self.bias = scale, bias
        self._local_search = local_search
        self._merge = merge

    def _default_search(self, ops):
        """Exact local search: the tensor-core path for query blocks (flags deferred), the streaming scan otherwise."""
        def search(q, k):
            n = self.rows.shape[0]
            if q.shape[0] >= ops.tc_min_batch(self.rows) and ops.batch_topk_supported(self.rows, k) and n >= 1024:
                return ops.exact_topk_batched(self.rows, q, k, self.scale, self.bias, row_base=self.row_base, defer=True)
            return ops.scan_topk(self.rows, q, k, self.scale, self.bias, row_base=self.row_base)

        def fixup(flags, idx, score, q, k):
            return ops.exact_topk_fixup(flags, idx, score, self.rows, q, k, self.scale, self.bias,
                                        row_base=self.row_base, stats=self.stats)
        return search, fixup

    def _gather_merge(self, idx: torch.Tensor, score: torch.Tensor, k: int, flags: Optional[torch.Tensor]):
        b = idx.shape[0]
        if self._packed:
            # product path: pack kernel -> ONE collective -> merge kernel (3 launches, no torch elementwise ops)
            payload = self._ops.pack_topk(idx.contiguous(), score.contiguous(), flags)
            gathered = torch.empty(self.world * b, 2 * k + 1, dtype=torch.int64, device=idx.device)
            dist.all_gather_into_tensor(gathered, payload, group=self.group)
            return self._ops.topk_merge_packed(gathered, self.world, b, k)
        # injected local_search / merge (CPU plumbing tests): same payload layout, torch ops
        payload = torch.empty(b, 2 * k + 1, dtype=torch.int64, device=idx.device)
        payload[:, :k] = idx
        payload[:, k:2 * k] = score.contiguous().view(torch.int32)
        payload[:, 2 * k] = flags if flags is not None else 0
        gathered = torch.empty(self.world * b, 2 * k + 1, dtype=torch.int64, device=idx.device)
        dist.all_gather_into_tensor(gathered, payload, group=self.group)
        gathered = gathered.view(self.world, b, 2 * k + 1)
        cat_idx = gathered[:, :, :k].permute(1, 0, 2).reshape(b, self.world * k).contiguous()
        cat_score = gathered[:, :, k:2 * k].to(torch.int32).view(torch.float32).permute(1, 0, 2) \
            .reshape(b, self.world * k).contiguous()
        out_score, out_idx = self._merge(cat_score, cat_idx, self.world, k, k)
        any_flag = gathered[:, :, 2 * k].sum(dim=0)               # identical on every rank
        return out_idx, out_score, any_flag

    def _flag_slot(self):
        """A pinned scalar + event for the flag count of one search, from a small ring: allocating pinned memory per
        search costs a cudaHostAlloc (tens of microseconds, serialised across the ranks of a box) on a 0.5 ms step.
        Up to 8 searches may be pending at once."""
        if not hasattr(self, "_slots"):
            self._slots = [(torch.empty(1, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(8)]
            self._slot_next = 0
        slot = self._slots[self._slot_next % len(self._slots)]
        self._slot_next += 1
        return slot

    def graphed(self, batch: int, k: int, allow_collective: bool = False) -> "GraphedSearch":
        """A CUDA-graph capture of `search_deferred` for a fixed (batch, k); see `GraphedSearch`."""
        if self.world > 1 and not allow_collective:
            raise RuntimeError("GraphedSearch with world > 1 captures an NCCL collective (unverified here); "
                               "pass allow_collective=True to try")
        return GraphedSearch(self, batch, k)

    def search(self, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries [B,d] identical on all ranks -> (global rows int64 [B,k], scores fp32 [B,k]) on all ranks."""
        return self.finalize(self.search_deferred(queries, k))

    def search_deferred(self, queries: torch.Tensor, k: int):
        """Enqueue one search (local top-k, pack, all-gather, merge, flag reduction + async copy of the flag count to
        pinned memory) and return a handle WITHOUT synchronising; `finalize(handle)` waits for exactly this search and
        runs the (rare) exact re-run.  A serving loop that finalises search i after enqueuing search i+1 keeps the GPU
        busy while the host prepares the next launches (bench.py does this)."""
        if queries.device != self.rows.device:
            queries = queries.to(self.rows.device, non_blocking=True)
        res = self._local_search(queries, k)
        idx, score = res[0], res[1]
        flags = res[2] if len(res) > 2 else None
        if self.world == 1:
            return {"idx": idx, "score": score, "flags": flags, "queries": queries, "k": k, "local": True}
        out_idx, out_score, any_flag = self._gather_merge(idx, score, k, flags)
        h = {"idx": out_idx, "score": out_score, "any_flag": any_flag, "queries": queries, "k": k, "local": False,
             "checked": flags is None}
        if flags is not None and queries.is_cuda:
            n_bad, ev = self._flag_slot()
            n_bad.copy_((any_flag != 0).sum().reshape(1), non_blocking=True)
            ev.record(torch.cuda.current_stream(queries.device))
            h["n_bad"], h["event"] = n_bad, ev
        return h

    def finalize(self, h) -> Tuple[torch.Tensor, torch.Tensor]:
        idx, score, queries, k = h["idx"], h["score"], h["queries"], h["k"]
        if h["local"]:
            if "event" in h:                              # graphed search: the flag count is already on its way
                h["event"].synchronize()
                if int(h["n_bad"][0]) == 0:
                    return idx, score
            if h["flags"] is not None:
                self._fixup(h["flags"], idx, score, queries, k)
            return idx, score
        if h["checked"]:
            return idx, score
        if "event" in h:
            h["event"].synchronize()                      # waits for this search only
            if int(h["n_bad"][0]) == 0:
                return idx, score
        # every rank sees the same flags, so the (rare) re-run below is entered by all ranks together
        bad = torch.nonzero(h["any_flag"], as_tuple=False).squeeze(-1)
        if bad.numel() > 0:
            qb = queries[bad].contiguous()
            fl = torch.ones(bad.numel(), dtype=torch.int32, device=idx.device)
            i2 = torch.empty(bad.numel(), k, dtype=idx.dtype, device=idx.device)
            s2 = torch.empty(bad.numel(), k, dtype=score.dtype, device=idx.device)
            self._fixup(fl, i2, s2, qb, k)
            i3, s3, _ = self._gather_merge(i2, s2, k, None)
            idx[bad] = i3
            score[bad] = s3
        return idx, score


class GraphedSearch:
    """One `ShardedBank` search of a fixed (batch, k) captured as a CUDA graph: local top-k kernels, pack, the NCCL
    all-gather and the merge replay as ONE launch, so a step is no longer paced by the host issuing ~7 launches (at 8
    GPUs a 1M-row bank leaves ~0.3 ms of kernel time per step, less than the launch gaps it replaces).

    Verified on one GPU (bench: 2.34 -> 2.28 ms per 1024-query step).  With world > 1 the capture includes the NCCL
    all-gather; on this image (torch 2.11, NCCL 2.28.9) a 2-rank capture hung, so `ShardedBank.graphed` refuses
    world > 1 unless `allow_collective=True`.

    The graph owns static buffers: `launch(queries)` copies the queries in, replays, and returns a handle for
    `ShardedBank.finalize`; results stay valid until the next `launch` of the SAME object - keep two objects and alternate
    them to have two searches in flight."""

    def __init__(self, bank: "ShardedBank", batch: int, k: int, warmup: int = 3):
        dev = bank.rows.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedSearch needs a CUDA bank")
        self.bank, self.k = bank, int(k)
        self.q = torch.zeros(batch, bank.rows.shape[1], device=dev, dtype=torch.float32)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # communicator, workspaces and kernel attributes exist before capture
                self._enqueue()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        lib = bank._ops._lib.load() if bank._ops is not None else None
        n0 = lib.aura_kernel_launches() if lib is not None else 0
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
            self.idx, self.score, self.flag, self.n_bad_dev = self._enqueue()
        self.kernels_per_replay = int(lib.aura_kernel_launches() - n0) if lib is not None else 0   # library kernels in the graph
        self._n_bad = torch.empty(1, dtype=torch.int64).pin_memory()

    def _enqueue(self):
        b = self.bank
        res = b._local_search(self.q, self.k)
        idx, score = res[0], res[1]
        flags = res[2] if len(res) > 2 else None
        if b.world > 1:
            idx, score, flags = b._gather_merge(idx, score, self.k, flags)
        if flags is None:
            flags = torch.zeros(idx.shape[0], dtype=torch.int32, device=idx.device)
        return idx, score, flags, (flags != 0).sum().reshape(1)

    def launch(self, queries: torch.Tensor):
        self.q.copy_(queries, non_blocking=True)
        self.graph.replay()
        self._n_bad.copy_(self.n_bad_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.q.device))
        if self.bank.world == 1:
            return {"idx": self.idx, "score": self.score, "flags": self.flag, "queries": self.q, "k": self.k,
                    "local": True, "n_bad": self._n_bad, "event": ev}
        return {"idx": self.idx, "score": self.score, "any_flag": self.flag, "queries": self.q, "k": self.k,
                "local": False, "checked": False, "n_bad": self._n_bad, "event": ev}


class ShardedIndex:
    """Row-sharded centroid index: every rank holds a `HippocampalFormation` over its own rows
    (global rows [row_base, row_base + memory_count)), the centroids are replicated and kept bit-identical on all
    ranks, every rank keeps its own CSR slice of every inverted list.

    rebuild : seed rows are fetched from their owners (sum all-reduce of a [k,d] block that only the owner fills),
              local assign -> local per-list fp64 sums + counts -> all-reduce -> identical new centroids everywhere
              -> local re-assign + local lists (hippocampal.py:345-377 distributed over ranks).
    search  : replicated queries, local coarse + fine search with global row ids, all-gather, k-way merge
              (`ShardedBank.search` plumbing).  nprobe means the same as in the single index, so recall is comparable.

    `all_reduce` / `all_gather` default to torch.distributed (NCCL on GPUs); they are injectable so that the
    arithmetic can be checked against a single index inside one process (tests/test_sharded_gpu.py).
    """

    def __init__(self, local, row_base: int, n_total: int, all_reduce: Optional[Callable] = None,
                 all_gather: Optional[Callable] = None, world: Optional[int] = None):
        self.local = local
        self.row_base = int(row_base)
        self.n_total = int(n_total)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        if all_reduce is None:
            def all_reduce(t):
                if self.world > 1:
                    dist.all_reduce(t)
                return t
        if all_gather is None:
            def all_gather(t):
                if self.world == 1:
                    return t.unsqueeze(0)
                out = torch.empty(self.world * t.shape[0], *t.shape[1:], dtype=t.dtype, device=t.device)
                dist.all_gather_into_tensor(out, t.contiguous())
                return out.view(self.world, *t.shape)
        self._all_reduce, self._all_gather = all_reduce, all_gather

    def rebuild_centroids(self, seed_rows_global: torch.Tensor) -> None:
        from . import ops
        hf = self.local
        m, dev = hf.memory_count, hf.device
        k = min(hf.centroids_k, self.n_total)
        rows_c = hf.centroids.shape[0]
        seeds = seed_rows_global.to(device=dev, dtype=torch.int64)[:k]
        mine = (seeds >= self.row_base) & (seeds < self.row_base + m)
        block = torch.zeros(k, hf.memory_features.shape[1], device=dev, dtype=torch.float32)
        if bool(mine.any()):
            tmp = torch.empty(int(mine.sum()), block.shape[1], device=dev, dtype=torch.float32)
            ops.kmeans_seed(hf.memory_features, (seeds[mine] - self.row_base).contiguous(), tmp)
            block[mine] = tmp
        self._all_reduce(block)                                   # every seed row has exactly one owner
        hf.centroids[:k] = block
        if k < rows_c:
            hf.centroids[k:].zero_()
        ops.kmeans_assign(hf.memory_features, m, hf.centroids, k, hf._cid, inv_norm=hf._inv_norm)
        ops.ivf_build_lists(hf._cid, m, rows_c, hf._list_offsets, hf._list_rows)
        sums = torch.empty(rows_c, block.shape[1], device=dev, dtype=torch.float64)
        counts = torch.empty(rows_c, device=dev, dtype=torch.int64)
        ops.kmeans_list_sums(hf.memory_features, hf._list_offsets, hf._list_rows, rows_c, sums, counts)
        self._all_reduce(sums)
        self._all_reduce(counts)
        ops.kmeans_finalize(sums, counts, k, hf.centroids)        # identical inputs -> identical centroids on all ranks
        ops.kmeans_assign(hf.memory_features, m, hf.centroids, k, hf._cid, hf.memory_metadata[:, 2], 4,
                          inv_norm=hf._inv_norm)
        ops.ivf_build_lists(hf._cid, m, rows_c, hf._list_offsets, hf._list_rows)
        local_counts = torch.empty(rows_c, device=dev, dtype=torch.float32)
        ops.ivf_list_counts(hf._list_offsets, rows_c, local_counts)
        self._all_reduce(local_counts)
        hf.centroid_counts = local_counts[:max(hf.centroids_k, 1)].clone()
        hf._lists_dirty = False
        hf._by_list_valid = False
        hf._index_ready = True

    def search(self, queries: torch.Tensor, k: int, exact: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """(global rows int64 [B,k], scores [B,k]) on every rank; centroid path unless `exact`."""
        from . import ops
        hf = self.local
        idx, score = hf.retrieve_batch(queries, k=k, force_exact=exact)
        idx = torch.where(idx >= 0, idx + self.row_base, idx)
        if idx.shape[1] < k:                                       # a shard with fewer than k rows
            pad = k - idx.shape[1]
            idx = torch.cat([idx, idx.new_full((idx.shape[0], pad), -1)], dim=1)
            score = torch.cat([score, score.new_full((score.shape[0], pad), float("-inf"))], dim=1)
        if self.world == 1:
            return idx, score
        b = idx.shape[0]
        payload = ops.pack_topk(idx.contiguous(), score.contiguous(), None)       # [B, 2k+1]
        gathered = self._all_gather(payload)                                       # [G, B, 2k+1], one collective
        g = gathered.shape[0]
        out_idx, out_score, _ = ops.topk_merge_packed(gathered.reshape(g * b, 2 * k + 1).contiguous(), g, b, k)
        return out_idx, out_score
