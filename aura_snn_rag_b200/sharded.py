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


class PeerGather:
    """The all-gather + merge of a sharded search over peer memory (NVLink / NVSwitch) instead of a library collective:
    one symmetric buffer per rank (`torch.distributed._symmetric_memory`: every rank maps every peer's buffer), into
    which `aura_pack_scatter` stores this rank's payload on EVERY rank, and which `aura_merge_gathered` merges once the
    flags of all ranks are up.  Two kernel launches per search, no host synchronisation, capturable in a CUDA graph.
    world == 1 (tests): a plain local buffer."""

    def __init__(self, world: int, rank: int, batch: int, k: int, device: torch.device, group=None):
        import ctypes
        from . import ops
        self.world, self.rank, self.batch, self.k = world, rank, batch, k
        n64 = (ops.peer_gather_buffer_bytes(world, batch, k) + 7) // 8
        if world > 1:
            import torch.distributed._symmetric_memory as symm
            self.buf = symm.empty(n64, dtype=torch.int64, device=device)
            self.buf.zero_()
            torch.cuda.synchronize(device)
            self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            ptrs = [int(p) for p in self.handle.buffer_ptrs]
            self.handle.barrier()                    # every buffer is zeroed before anyone stores into it
        else:
            self.buf = torch.zeros(n64, dtype=torch.int64, device=device)
            self.handle = None
            ptrs = [self.buf.data_ptr()]
        self.ptrs = (ctypes.c_void_p * world)(*ptrs)
        self.counters = torch.zeros(4, dtype=torch.int32, device=device)
        self._ops = ops

    def gather_merge(self, idx, score, flags, id_map=None, id_base: int = 0):
        self._ops.pack_scatter(idx.contiguous(), score.contiguous(), flags, self.ptrs, self.rank, self.world, self.counters,
                               id_map=id_map, id_base=id_base)
        return self._ops.merge_gathered(self.buf, self.world, idx.shape[0], idx.shape[1], self.counters)


class ShardedBank:
    def __init__(self, local_rows: torch.Tensor, row_base: int, group=None,
                 scale: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 stats: Optional[dict] = None, shadow=None, score_unit: float = 1.0, peer_gather: bool = False):
        self.rows = local_rows
        # exchange the local top-k blocks through peer memory (PeerGather) instead of an NCCL all-gather; decided
        # collectively at the first search, falls back to NCCL if the symmetric allocation is not available
        self.peer_gather = bool(peer_gather)
        self._peers = {}
        self.shadow = shadow                  # ops.Bf16Shadow of an fp32 shard: the shortlist pass reads it (ops.batch_topk)
        self.score_unit = float(score_unit)   # max |scale_r| * ||r|| (1 for pure cosine): unit of the certification bound
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
                sh = self.shadow if k <= ops.TC_SHADOW_MAX_K else None
                return ops.exact_topk_batched(self.rows, q, k, self.scale, self.bias, row_base=self.row_base, defer=True,
                                              eps=self.score_unit * (1.0 if sh is not None else ops.TC_EPS_COS), shadow=sh)
            return ops.scan_topk(self.rows, q, k, self.scale, self.bias, row_base=self.row_base)

        def fixup(flags, idx, score, q, k):
            return ops.exact_topk_fixup(flags, idx, score, self.rows, q, k, self.scale, self.bias,
                                        row_base=self.row_base, stats=self.stats)
        return search, fixup

    def _gather_merge(self, idx: torch.Tensor, score: torch.Tensor, k: int, flags: Optional[torch.Tensor]):
        b = idx.shape[0]
        if self._packed and self.peer_gather:
            pg = self._peer_gather_for(b, k, idx.device)
            if pg is not None:
                return pg.gather_merge(idx, score, flags)
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

    def _peer_gather_for(self, b: int, k: int, device) -> Optional[PeerGather]:
        """The PeerGather of this (batch, k), created collectively on first use; None (for good) if any rank could not
        set up symmetric memory."""
        key = (b, k)
        if key in self._peers:
            return self._peers[key]
        pg = None
        try:
            pg = PeerGather(self.world, self.rank, b, k, device, self.group)
            ok = torch.ones(1, device=device)
        except Exception as e:                                  # noqa: BLE001 - agreed on collectively below
            import sys
            sys.stderr.write(f"[aura] peer gather unavailable on rank {self.rank} ({e}); using the NCCL all-gather\n")
            ok = torch.zeros(1, device=device)
        if self.world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if float(ok) < 1.0:
            pg = None
            self.peer_gather = False
        self._peers[key] = pg
        return pg

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
        if self.world > 1 and not (allow_collective or self.peer_gather):
            raise RuntimeError("GraphedSearch with world > 1 captures an NCCL collective (unverified here); "
                               "use peer_gather=True (exchange over peer memory) or pass allow_collective=True to try")
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
        if bank._ops is not None:
            # the graph bakes in the raw pointers of this stream's scratch buffers: they must outlive the graph even if
            # a later, larger request on the same stream handle re-allocates them (torch recycles stream handles)
            bank._ops.pin_workspaces(dev, side.cuda_stream)
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
    """Row-sharded centroid index: every rank holds a `HippocampalFormation` over its own rows, the
    centroids are replicated and kept bit-identical on all ranks, every rank keeps its own CSR slice of every inverted
    list.  Local row r of this rank is global memory `gid[r]`: initially the contiguous range
    [row_base, row_base + memory_count); online writes append global ids n_total, n_total + 1, ... to the rank
    `id % world` (the owner rule of SURVEY 8e), so the table stays ascending.

    rebuild : seed rows are fetched from their owners (sum all-reduce of a [k,d] block that only the owner fills),
              local assign -> local per-list fp64 sums + counts -> all-reduce -> identical new centroids everywhere
              -> local re-assign + local lists (hippocampal.py:345-377 distributed over ranks).
    write   : `create_episodic_memories(features)` with the SAME feature block on every rank: every rank runs the
              sequential nearest-centroid + running-mean update (hippocampal.py:218-230) over all rows, in order, on a
              staging copy - the replicas of the centroids and counts therefore never diverge - and stores only the
              rows it owns.
    search  : replicated queries, local coarse + fine search, ids mapped to global inside the pack kernel, ONE
              all-gather, k-way merge.  nprobe means the same as in the single index, so recall is comparable.  A shard
              whose slices of the probed lists are empty contributes nothing; the reference's "no candidates -> all
              rows" rule (hippocampal.py:269-270) is applied to the merged result.
    map     : `build_cognitive_map`: one all-gather of the rows, each rank scores its own rows against the whole bank.

    `all_reduce` / `all_gather` default to torch.distributed (NCCL on GPUs); they are injectable so that the
    arithmetic can be checked against a single index inside one process (tests/test_sharded_gpu.py).
    """

    def __init__(self, local, row_base: int, n_total: int, all_reduce: Optional[Callable] = None,
                 all_gather: Optional[Callable] = None, world: Optional[int] = None, rank: Optional[int] = None):
        self.local = local
        self.row_base = int(row_base)
        self.n_total = int(n_total)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
        # local row -> global id; None while the shard is still the contiguous range it was built from
        self.gid: Optional[torch.Tensor] = None
        self.update_interval = int(getattr(local, "centroids_update_interval", 512))
        local.centroids_update_interval = 1 << 62          # rebuild points are decided here, on the global count
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

    # ------------------------------------------------------------------ ids
    def _gid_table(self) -> torch.Tensor:
        hf = self.local
        if self.gid is None:
            self.gid = torch.arange(self.row_base, self.row_base + hf.max_memories, device=hf.device, dtype=torch.int64)
            self._gid_contiguous = hf.memory_count
        return self.gid

    def _local_rows_of(self, global_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(mask of the ids this rank owns, their local rows)."""
        m = self.local.memory_count
        if self.gid is None:
            mine = (global_ids >= self.row_base) & (global_ids < self.row_base + m)
            return mine, (global_ids[mine] - self.row_base)
        table = self.gid[:m]
        pos = torch.searchsorted(table, global_ids).clamp_(max=max(m - 1, 0))
        mine = table[pos] == global_ids if m > 0 else torch.zeros_like(global_ids, dtype=torch.bool)
        return mine, pos[mine]

    # ------------------------------------------------------------------ build
    def rebuild_centroids(self, seed_rows_global: torch.Tensor) -> None:
        from . import ops
        hf = self.local
        m, dev = hf.memory_count, hf.device
        k = min(hf.centroids_k, self.n_total)
        rows_c = hf.centroids.shape[0]
        seeds = seed_rows_global.to(device=dev, dtype=torch.int64)[:k]
        mine, local_rows = self._local_rows_of(seeds)
        block = torch.zeros(k, hf.memory_features.shape[1], device=dev, dtype=torch.float32)
        if bool(mine.any()):
            tmp = torch.empty(int(mine.sum()), block.shape[1], device=dev, dtype=torch.float32)
            ops.kmeans_seed(hf.memory_features, local_rows.contiguous(), tmp)
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

    # ------------------------------------------------------------------ online writes
    def create_episodic_memories(self, features: torch.Tensor, seed_rows_fn: Optional[Callable] = None) -> Tuple[int, int]:
        """Append n memories (global ids n_total .. n_total + n); `features` [n,d] must be identical on every rank.
        Same semantics as n sequential `create_episodic_memory` calls on a single index, including the rebuild points
        (every `update_interval` memories once the global count exceeds centroids_k, hippocampal.py:242-243);
        `seed_rows_fn(count)` supplies the global seed rows of a triggered rebuild (default: randperm of the global
        count, seeded by the count so that every rank draws the same rows).  Returns the global id range written."""
        from . import ops
        hf = self.local
        dev = hf.device
        feats = hf._to_dev_f32(features).detach()
        if feats.dim() == 1:
            feats = feats.unsqueeze(0)
        feats = feats.contiguous()
        n, d = feats.shape
        first = self.n_total
        gid = self._gid_table()
        done = 0
        while done < n:
            nxt = (self.n_total // self.update_interval + 1) * self.update_interval
            step = min(n - done, nxt - self.n_total)
            blk = feats[done:done + step]
            # staging bank: the rows exactly as the bank stores them (dtype conversion, inverse norms) + their metadata
            st_rows = torch.empty(step, d, device=dev, dtype=hf.memory_features.dtype)
            st_meta = torch.empty(step, 4, device=dev, dtype=torch.float32)
            st_loc = torch.empty(step, hf.memory_locations.shape[1], device=dev, dtype=torch.float32)
            st_inv = torch.empty(step, device=dev, dtype=torch.float32)
            st_cid = torch.full((step,), -1, device=dev, dtype=torch.int32)
            ops.bank_write(st_rows, 0, blk, st_meta, st_inv, hf_time(hf), st_loc, hf.current_location.contiguous())
            if hf.use_centroid_index and hf._index_ready:
                live = min(hf.centroids_k, hf.centroids.shape[0])
                # every rank updates the replicated centroids with EVERY row, in order: replicas stay bit-identical
                ops.online_assign(st_rows, 0, step, hf.centroids, live, hf.centroid_counts, st_cid, st_meta[:, 2], 4)
            ids = torch.arange(self.n_total, self.n_total + step, device=dev, dtype=torch.int64)
            own = (ids % self.world) == self.rank
            n_own = int(own.sum())
            m = hf.memory_count
            if m + n_own > hf.max_memories:
                raise ValueError(f"shard {self.rank}: {n_own} new rows do not fit ({m} of {hf.max_memories} used)")
            if n_own:
                hf.memory_features[m:m + n_own] = st_rows[own]
                hf.memory_metadata[m:m + n_own] = st_meta[own]
                hf.memory_locations[m:m + n_own] = st_loc[own]
                hf._inv_norm[m:m + n_own] = st_inv[own]
                hf._cid[m:m + n_own] = st_cid[own]
                gid[m:m + n_own] = ids[own]
                hf.memory_count = m + n_own
                hf._lists_dirty = True
                hf._version += 1
            self.n_total += step
            done += step
            if hf.use_centroid_index and self.n_total % self.update_interval == 0 and self.n_total > hf.centroids_k:
                if seed_rows_fn is not None:
                    seeds = seed_rows_fn(self.n_total)
                else:
                    g = torch.Generator().manual_seed(self.n_total)
                    seeds = torch.randperm(self.n_total, generator=g)[:hf.centroids_k]
                self.rebuild_centroids(seeds)
        return first, first + n

    # ------------------------------------------------------------------ query
    def search(self, queries: torch.Tensor, k: int, exact: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """(global ids int64 [B,k], scores [B,k]) on every rank; centroid path unless `exact`."""
        from . import ops
        hf = self.local
        idx, score = hf.retrieve_batch(queries, k=k, force_exact=exact, allow_empty=self.world > 1)
        if idx.shape[1] < k:                                       # a shard with fewer than k rows
            pad = k - idx.shape[1]
            idx = torch.cat([idx, idx.new_full((idx.shape[0], pad), -1)], dim=1)
            score = torch.cat([score, score.new_full((score.shape[0], pad), float("-inf"))], dim=1)
        b = idx.shape[0]
        payload = ops.pack_topk(idx.contiguous(), score.contiguous(), None, id_map=self.gid,
                                id_base=self.row_base)                             # [B, 2k+1], global ids
        gathered = self._all_gather(payload)                                       # [G, B, 2k+1], one collective
        g = gathered.shape[0]
        out_idx, out_score, _ = ops.topk_merge_packed(gathered.reshape(g * b, 2 * k + 1).contiguous(), g, b, k)
        if not exact and self.world > 1 and hf._centroid_path():
            # hippocampal.py:269-270 on the merged result: no candidate on ANY shard -> scan all rows (all ranks agree)
            empty = torch.nonzero(out_idx[:, 0] < 0, as_tuple=False).squeeze(-1)
            if empty.numel() > 0:
                q = hf._queries(queries)
                i2, s2 = self.search(q[empty].contiguous(), k, exact=True)
                out_idx[empty] = i2
                out_score[empty] = s2
        return out_idx, out_score

    # ------------------------------------------------------------------ cognitive map
    def build_cognitive_map(self, k: int = 32) -> Tuple[torch.Tensor, torch.Tensor]:
        """Sharded all-pairs map: ONE all-gather of the rows (padded to the largest shard), then every rank
        scores its own rows (the A block) against the whole bank with the top-k fused (`aura_allpairs_topk`); no result
        exchange.  Returns (global neighbour ids int64 [m_local, k], cosine fp32 [m_local, k]) for this rank's rows."""
        from . import ops
        hf = self.local
        dev, m, d = hf.device, hf.memory_count, hf.memory_features.shape[1]
        counts = self._all_gather(torch.tensor([m], device=dev, dtype=torch.int64)).reshape(-1)
        counts_l = counts.tolist()
        m_max = max(counts_l)
        pad = torch.zeros(m_max, d, device=dev, dtype=hf.memory_features.dtype)
        pad[:m] = hf.memory_features[:m]
        everyone = self._all_gather(pad)                                         # [G, m_max, d]
        ids = torch.full((m_max,), -1, device=dev, dtype=torch.int64)
        ids[:m] = self._gid_table()[:m] if self.gid is not None else torch.arange(self.row_base, self.row_base + m, device=dev)
        all_ids = self._all_gather(ids)                                          # [G, m_max]
        bank = torch.cat([everyone[r, :c] for r, c in enumerate(counts_l)], 0).contiguous()
        gids = torch.cat([all_ids[r, :c] for r, c in enumerate(counts_l)], 0)
        a_first = sum(counts_l[:self.rank])
        kk = min(int(k), bank.shape[0] - 1)
        nbr, sim = ops.allpairs_topk(bank, kk, ops.row_inv_norms(bank), a_first=a_first, n_a=m)
        return torch.where(nbr >= 0, gids[nbr.clamp(min=0)], nbr), sim


def hf_time(hf) -> float:
    """The wall clock through the module `hippocampal` imported (tests freeze it there)."""
    from . import hippocampal as _h
    return _h.time.time()
