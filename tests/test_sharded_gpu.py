"""Sharded centroid index on one GPU: 3 shards in one process with emulated collectives must reproduce the
single-index build (centroids, assignments, counts) and its query results; the NCCL plumbing itself is
exercised by `bench.py --gpus N` and the gloo test."""
import types

import numpy as np
import pytest
import torch

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

    # emulate the collectives: run every rank up to the collective, reduce, continue (lock-step driver)
    import threading
    barrier = threading.Barrier(world)
    slots = {}
    lock = threading.Lock()

    def make_collectives(rank):
        state = {"n": 0}

        def all_reduce(t):
            key = ("r", state["n"]); state["n"] += 1
            with lock:
                slots.setdefault(key, []).append(t)
            barrier.wait()
            if rank == 0:
                total = torch.stack([x.double() if x.is_floating_point() else x for x in slots[key]]).sum(0)
                for x in slots[key]:
                    x.copy_(total.to(x.dtype))
            barrier.wait()
            return t

        def all_gather(t):
            key = ("g", state["n"]); state["n"] += 1
            with lock:
                slots.setdefault(key, {})[rank] = t
            barrier.wait()
            out = torch.stack([slots[key][r] for r in range(world)])
            barrier.wait()
            return out
        return all_reduce, all_gather

    results = [None] * world

    def run(rank):
        torch.cuda.set_device(0)
        hf, lo = shards[rank]
        ar, ag = make_collectives(rank)
        si = ShardedIndex(hf, lo, n, all_reduce=ar, all_gather=ag, world=world)
        si.rebuild_centroids(seeds)
        results[rank] = si.search(q.cuda(), k)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    torch.cuda.synchronize()
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
