"""Sharded centroid index on one GPU: 3 shards in one process with emulated collectives must reproduce the
single-index build (centroids, assignments, counts) and its query results; the NCCL plumbing itself is
exercised by `bench.py --gpus N` and the gloo test."""
import types

import numpy as np
import pytest
import torch

from shard_emu import Emulator

pytestmark = pytest.mark.gpu


def test_sharded_index_equals_single_index(monkeypatch):
    import aura_snn_rag_b200.hippocampal as hmod
    from aura_snn_rag_b200.sharded import ShardedIndex, shard_range
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: 1.79e9))
    g = torch.Generator().manual_seed(17)
    n, d, c, p, k, world = 9000, 64, 48, 6, 10, 3
    centres = torch.randn(24, d, generator=g)
    rows = centres[torch.randint(0, 24, (n,), generator=g)] + 0.4 * torch.randn(n, d, generator=g)
    seeds = torch.randperm(n, generator=g)[:c]
    q = rows[torch.randint(0, n, (40,), generator=g)] + 0.1 * torch.randn(40, d, generator=g)

    def make(m):
        hf = hmod.HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=m, feature_dim=d,
                                       centroids_k=c, centroid_rows=c + 8, nprobe=p, track_ids=False)
        hf.centroids_update_interval = 1 << 40
        return hf

    single = make(n)
    single.create_episodic_memories(rows)
    single.rebuild_centroids(seed_rows=seeds)
    ref_idx, ref_sc = single.retrieve_batch(q, k)

    shards = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        hf = make(hi - lo)
        hf.create_episodic_memories(rows[lo:hi])
        shards.append((hf, lo))

    def run(rank, ar, ag):
        hf, lo = shards[rank]
        si = ShardedIndex(hf, lo, n, all_reduce=ar, all_gather=ag, world=world, rank=rank)
        si.rebuild_centroids(seeds)
        return si.search(q.cuda(), k)

    results = Emulator(world).run(run)
    for hf, lo in shards:
        np.testing.assert_allclose(hf.centroids.cpu().numpy(), single.centroids.cpu().numpy(), rtol=1e-5, atol=1e-6)
        m = hf.memory_count
        assert torch.equal(hf._cid[:m].cpu(), single._cid[lo:lo + m].cpu())
        assert torch.equal(hf.centroid_counts.cpu(), single.centroid_counts.cpu())
    for r in range(world):
        idx, sc = results[r]
        assert torch.equal(idx.cpu(), ref_idx.cpu())
        np.testing.assert_allclose(sc.cpu().numpy(), ref_sc.cpu().numpy(), rtol=1e-6)


def test_pack_and_merge_packed_equal_the_reference_merge():
    from aura_snn_rag_b200 import ops
    from oracle.hippo_oracle import merge_topk
    g = torch.Generator().manual_seed(3)
    G, B, k = 5, 37, 10
    scores = torch.randn(G, B, k, generator=g)
    scores[1, :, 3] = scores[3, :, 7]                      # cross-rank score ties
    ids = torch.randperm(10 ** 7, generator=g)[: G * B * k].reshape(G, B, k) + 3 * 10 ** 9
    ids[2, 5, 4:] = -1                                     # a rank with fewer than k results
    flags = torch.zeros(G, B, dtype=torch.int32); flags[4, 11] = 1
    payloads = [ops.pack_topk(ids[r].cuda(), scores[r].cuda(), flags[r].cuda()) for r in range(G)]
    gathered = torch.cat(payloads, 0)
    oi, os_, fl = ops.topk_merge_packed(gathered, G, B, k)
    s2 = scores.clone(); s2[ids < 0] = -float("inf")
    ref_s, ref_i = merge_topk(s2.permute(1, 0, 2).reshape(B, G * k), ids.permute(1, 0, 2).reshape(B, G * k), k)
    assert torch.equal(oi.cpu(), ref_i) and torch.equal(os_.cpu(), ref_s)
    assert fl.cpu().tolist() == [1 if b == 11 else 0 for b in range(B)]


def test_graphed_search_equals_eager_search():
    """`ShardedBank.graphed` (the whole search captured as one CUDA graph, static buffers) returns exactly what the eager
    search returns, for several query batches replayed through two alternating graphs."""
    from aura_snn_rag_b200 import ops
    from aura_snn_rag_b200.sharded import ShardedBank
    g = torch.Generator().manual_seed(5)
    rows = torch.randn(20000, 128, generator=g).cuda()
    bank = ShardedBank(rows, 1000, scale=ops.row_inv_norms(rows))
    graphs = [bank.graphed(64, 10) for _ in range(2)]
    assert graphs[0].kernels_per_replay >= 3          # normalise + tensor-core scoring + finish are inside the graph
    pending = None
    for i in range(5):
        q = torch.randn(64, 128, generator=g).cuda()
        ref_idx, ref_sc = bank.search(q, 10)
        h = graphs[i % 2].launch(q)
        if pending is not None:                       # results of the other graph are still intact
            idx_p, sc_p = bank.finalize(pending[0])
            assert torch.equal(idx_p, pending[1]) and torch.equal(sc_p, pending[2])
        idx, sc = bank.finalize(h)
        assert torch.equal(idx, ref_idx) and torch.equal(sc, ref_sc)
        assert int(idx.min()) >= 1000                 # global row ids
        pending = (h, ref_idx, ref_sc)


def _clustered(n, d, n_centres, sigma, seed, sort_by_cluster=False):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(n_centres, d, generator=g)
    which = torch.randint(0, n_centres, (n,), generator=g)
    if sort_by_cluster:
        which, _ = torch.sort(which)
    return centres[which] + sigma * torch.randn(n, d, generator=g), which, g


def test_shard_without_local_candidates_returns_nothing(monkeypatch):
    """Contiguous row shards of a bank whose rows arrive cluster by cluster: for most queries two of the three shards hold
    none of the probed lists' rows.  Such a shard must contribute NOTHING (AURA_IVF_EMPTY_OK) - the reference's 'no
    candidates -> all rows' rule (hippocampal.py:269-270) applies to the merged candidate set, not per shard - so the
    sharded result equals the single index bit for bit."""
    import aura_snn_rag_b200.hippocampal as hmod
    from aura_snn_rag_b200 import ops
    from aura_snn_rag_b200.sharded import ShardedIndex, shard_range
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: 1.79e9))
    n, d, c, p, k, world = 12000, 64, 48, 4, 10, 3
    rows, which, g = _clustered(n, d, 24, 0.25, 23, sort_by_cluster=True)
    seeds = torch.randperm(n, generator=g)[:c]
    q = rows[torch.randint(0, n, (96,), generator=g)] + 0.05 * torch.randn(96, d, generator=g)

    def make(m):
        hf = hmod.HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=m, feature_dim=d,
                                       centroids_k=c, centroid_rows=c + 8, nprobe=p, track_ids=False)
        hf.centroids_update_interval = 1 << 40
        return hf

    single = make(n)
    single.create_episodic_memories(rows)
    single.rebuild_centroids(seed_rows=seeds)
    ref = {b: single.retrieve_batch(q[:b], k) for b in (96, 5)}          # batched (tensor-core) and per-query paths
    shards = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        hf = make(hi - lo)
        hf.create_episodic_memories(rows[lo:hi])
        shards.append((hf, lo))

    def run(rank, ar, ag):
        hf, lo = shards[rank]
        si = ShardedIndex(hf, lo, n, all_reduce=ar, all_gather=ag, world=world, rank=rank)
        si.rebuild_centroids(seeds)
        return {b: si.search(q[:b].cuda(), k) for b in (96, 5)}

    results = Emulator(world).run(run)
    for r in range(world):
        for b in (96, 5):
            assert torch.equal(results[r][b][0].cpu(), ref[b][0].cpu())
            np.testing.assert_allclose(results[r][b][1].cpu().numpy(), ref[b][1].cpu().numpy(), rtol=1e-6)
    # the situation really occurred: on the last shard most of these queries have no local candidate at all
    hf2 = shards[2][0]
    for b in (96, 5):
        idx_e, sc_e = hf2.retrieve_batch(q[:b], k, allow_empty=True)
        empty = idx_e[:, 0] < 0
        assert float(empty.float().mean()) > 0.3
        assert bool((idx_e[empty] == -1).all()) and bool(torch.isinf(sc_e[empty]).all())
        idx_f, _ = hf2.retrieve_batch(q[:b], k)                            # single-index rule: those queries scan every row
        assert bool((idx_f[empty, 0] >= 0).all())
    # the reference's rule on the MERGED result: a query whose probed lists are empty everywhere scans all rows
    lists_empty_everywhere = torch.nonzero(single.centroid_counts[:c] == 0).squeeze(-1)
    if lists_empty_everywhere.numel() >= p:
        pass  # (not reachable with one Lloyd step on this data; covered by the ops-level case below)
    off = torch.zeros(c + 9, dtype=torch.int32, device="cuda")             # an index whose lists are all empty
    lr = torch.zeros(n, dtype=torch.int32, device="cuda")
    sc_, bi_ = single._row_terms(None)
    i0, s0 = ops.ivf_search(single.memory_features, n, q[:3].cuda(), single.centroids, p, off, lr, k, sc_, bi_, allow_empty=True)
    assert bool((i0 == -1).all())
    i1, s1 = ops.ivf_search(single.memory_features, n, q[:3].cuda(), single.centroids, p, off, lr, k, sc_, bi_)
    ie, se = ops.scan_topk(single.memory_features, q[:3].cuda(), k, sc_, bi_, n_rows=n)
    assert torch.equal(i1, ie) and torch.equal(s1, se)                     # hippocampal.py:269-270


def test_sharded_writes_keep_replicas_identical_and_match_single_index(monkeypatch):
    """Interleaved online writes and rebuilds through `ShardedIndex.create_episodic_memories`: rows go to rank id % world,
    every rank runs the sequential centroid update over all rows, so the centroid replicas stay bit-identical across
    ranks and the index (assignments, counts, query results) equals a single index fed the same rows."""
    import aura_snn_rag_b200.hippocampal as hmod
    from aura_snn_rag_b200.sharded import ShardedIndex, shard_range
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: 1.79e9))
    n0, n_new, d, c, p, k, world = 6000, 2600, 64, 32, 5, 10, 3
    rows, _, g = _clustered(n0 + n_new, d, 20, 0.35, 41)
    interval = 1024
    seeds0 = torch.randperm(n0, generator=g)[:c]
    seed_for = {cnt: torch.randperm(cnt, generator=torch.Generator().manual_seed(cnt))[:c]
                for cnt in range(interval, n0 + n_new + 1, interval)}
    q = rows[torch.randint(0, n0 + n_new, (80,), generator=g)] + 0.1 * torch.randn(80, d, generator=g)

    def make(m):
        hf = hmod.HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=m, feature_dim=d,
                                       centroids_k=c, centroid_rows=c, nprobe=p, track_ids=False)
        hf.centroids_update_interval = 1 << 40
        return hf

    single = make(n0 + n_new)
    single.create_episodic_memories(rows[:n0])
    single.rebuild_centroids(seed_rows=seeds0)
    single.centroids_update_interval = interval
    orig = single.rebuild_centroids
    single.rebuild_centroids = lambda seed_rows=None: orig(seed_rows=seed_for[single.memory_count])
    for a, b in ((n0, n0 + 700), (n0 + 700, n0 + 701), (n0 + 701, n0 + n_new)):       # three write bursts
        single.create_episodic_memories(rows[a:b])
    ref_idx, ref_sc = single.retrieve_batch(q, k)
    ref_ex, _ = single.retrieve_batch(q, k, force_exact=True)

    shards = []
    for r in range(world):
        lo, hi = shard_range(n0, r, world)
        hf = make(hi - lo + n_new)                                   # room for the rows this rank will own
        hf.create_episodic_memories(rows[lo:hi])
        shards.append((hf, lo))

    def run(rank, ar, ag):
        hf, lo = shards[rank]
        hf.centroids_update_interval = interval
        si = ShardedIndex(hf, lo, n0, all_reduce=ar, all_gather=ag, world=world, rank=rank)
        si.rebuild_centroids(seeds0)
        for a, b in ((n0, n0 + 700), (n0 + 700, n0 + 701), (n0 + 701, n0 + n_new)):
            si.create_episodic_memories(rows[a:b].cuda(), seed_rows_fn=lambda cnt: seed_for[cnt])
        return si, si.search(q.cuda(), k), si.search(q.cuda(), k, exact=True)

    results = Emulator(world).run(run)
    cents = [shards[r][0].centroids.cpu() for r in range(world)]
    for r in range(1, world):
        assert torch.equal(cents[r], cents[0])                       # replicas never diverge
        assert torch.equal(shards[r][0].centroid_counts.cpu(), shards[0][0].centroid_counts.cpu())
    np.testing.assert_allclose(cents[0].numpy(), single.centroids.cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert torch.equal(shards[0][0].centroid_counts.cpu(), single.centroid_counts.cpu())
    total = 0
    for r in range(world):
        si = results[r][0]
        hf = si.local
        m = hf.memory_count
        total += m
        gid = si.gid[:m].cpu() if si.gid is not None else torch.arange(si.row_base, si.row_base + m)
        assert torch.equal(hf._cid[:m].cpu(), single._cid[gid.cuda()].cpu())             # same list for every memory
        assert torch.equal(hf.memory_features[:m].cpu(), single.memory_features[gid.cuda()].cpu())
        idx, sc = results[r][1]
        assert torch.equal(idx.cpu(), ref_idx.cpu())
        np.testing.assert_allclose(sc.cpu().numpy(), ref_sc.cpu().numpy(), rtol=1e-6)
        assert torch.equal(results[r][2][0].cpu(), ref_ex.cpu())
    assert total == n0 + n_new


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_sharded_cognitive_map_equals_single_gpu_map(dt, monkeypatch):
    """`ShardedIndex.build_cognitive_map`: all-gather of the rows, per-rank A block -> the rows of the single-GPU map."""
    import aura_snn_rag_b200.hippocampal as hmod
    from aura_snn_rag_b200.sharded import ShardedIndex, shard_range
    n, d, k, world = 5000, 128, 16, 3
    rows, _, g = _clustered(n, d, 40, 0.6, 57)

    def make(m):
        return hmod.HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=m, feature_dim=d,
                                         use_centroid_index=False, bank_dtype=dt, track_ids=False)

    single = make(n)
    single.create_episodic_memories(rows)
    ref_nbr, ref_sim = single.build_cognitive_map(k)
    shards = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        hf = make(hi - lo)
        hf.create_episodic_memories(rows[lo:hi])
        shards.append((hf, lo))

    def run(rank, ar, ag):
        hf, lo = shards[rank]
        return ShardedIndex(hf, lo, n, all_reduce=ar, all_gather=ag, world=world, rank=rank).build_cognitive_map(k)

    results = Emulator(world).run(run)
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        nbr, sim = results[r]
        assert torch.equal(nbr.cpu(), ref_nbr[lo:hi].cpu())
        assert torch.equal(sim.cpu(), ref_sim[lo:hi].cpu())


def test_peer_gather_single_rank_equals_packed_merge():
    """`PeerGather` (pack + scatter into the gather buffer + flag, device-side wait + merge) with one rank: same result as
    pack -> merge_packed, over several steps (slot parity, step counters) and inside a CUDA graph."""
    from aura_snn_rag_b200 import ops
    from aura_snn_rag_b200.sharded import PeerGather
    g = torch.Generator().manual_seed(9)
    B, k = 33, 10
    pg = PeerGather(1, 0, B, k, torch.device("cuda:0"))
    for step in range(5):
        sc = torch.randn(B, k, generator=g).sort(dim=1, descending=True).values.cuda()
        ids = torch.randint(0, 10 ** 6, (B, k), generator=g).cuda()
        fl = (torch.rand(B, generator=g) < 0.1).int().cuda()
        ref_i, ref_s, ref_f = ops.topk_merge_packed(ops.pack_topk(ids, sc, fl, id_base=7), 1, B, k)
        i, s_, f = pg.gather_merge(ids, sc, fl, id_base=7)
        assert torch.equal(i, ref_i) and torch.equal(s_, ref_s) and torch.equal(f, ref_f)
    assert pg.counters.tolist() == [5, 0, 5, 0]
    sc = torch.randn(B, k, generator=g).sort(dim=1, descending=True).values.cuda()
    ids = torch.randint(0, 10 ** 6, (B, k), generator=g).cuda()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        out = pg.gather_merge(ids, sc, None)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    ref_i, ref_s, _ = ops.topk_merge_packed(ops.pack_topk(ids, sc, None), 1, B, k)
    assert torch.equal(out[0], ref_i) and torch.equal(out[1], ref_s) and pg.counters.tolist() == [8, 0, 8, 0]
