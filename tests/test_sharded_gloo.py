"""Row-sharded search plumbing on CPU: 2 processes, gloo backend.  The local search and the merge are
injected (CPU oracle helpers) so only the shard ranges, global row ids, all-gather layout and merge
order - the host logic of aura_snn_rag_b200/sharded.py - are under test; the CUDA kernels that fill
those roles in production are covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, d, b, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from aura_snn_rag_b200.sharded import ShardedBank, shard_range
    from oracle.hippo_oracle import exact_cosine_topk, merge_topk
    g = torch.Generator().manual_seed(5)
    bank = torch.randn(n, d, generator=g)
    bank[n // 2 + 3] = bank[7]                      # a cross-shard exact tie: lower global row must win
    q = torch.randn(b, d, generator=g)
    q[0] = bank[7]
    lo, hi = shard_range(n, rank, world)

    def local_search(queries, kk):
        idx, sc = exact_cosine_topk(bank[lo:hi], queries, kk)
        return idx + lo, sc

    def merge(scores, ids, n_lists, k_in, k_out):
        return merge_topk(scores, ids, k_out)

    sb = ShardedBank(bank[lo:hi], lo, local_search=local_search, merge=merge)
    idx, sc = sb.search(q, k)
    if rank == 0:
        ref_i, ref_s = exact_cosine_topk(bank, q, k)
        torch.save({"idx": idx, "sc": sc, "ref_i": ref_i, "ref_s": ref_s}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_search_two_ranks_gloo(tmp_path):
    from aura_snn_rag_b200.sharded import shard_range
    assert [shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    out = str(tmp_path / "res.pt")
    n, d, b, k = 1001, 32, 6, 5
    mp.spawn(_worker, args=(2, _free_port(), n, d, b, k, out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["sc"], r["ref_s"], rtol=1e-5, atol=1e-6)
    same = (r["idx"] == r["ref_i"])
    assert same[1:].all()
    # query 0 ties rows 7 and n//2+3 exactly: the merge must list the lower global row first
    assert r["idx"][0, 0].item() == 7 and r["idx"][0, 1].item() == n // 2 + 3


def test_graphed_search_needs_a_cuda_bank():
    """`ShardedBank.graphed` captures CUDA kernels: with a CPU bank (injected search / merge) it must refuse loudly."""
    import pytest
    import torch
    from aura_snn_rag_b200.sharded import ShardedBank
    rows = torch.randn(64, 8)
    bank = ShardedBank(rows, 0, scale=torch.ones(64), local_search=lambda q, k: (None, None), merge=lambda *a: None)
    with pytest.raises(RuntimeError):
        bank.graphed(4, 2)


def _index_worker(rank, world, port, q):
    """Host-side logic of ShardedIndex on 2 gloo ranks: the default collectives, the global-id table that online writes
    extend (owner = id % world), and the ownership look-up the sharded rebuild uses for its seed rows."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        from aura_snn_rag_b200.sharded import ShardedIndex, shard_range
        n0 = 10
        lo, hi = shard_range(n0, rank, world)
        local = types.SimpleNamespace(memory_count=hi - lo, max_memories=32, device=torch.device("cpu"),
                                      centroids_update_interval=512)
        si = ShardedIndex(local, lo, n0)
        assert si.world == world and si.rank == rank and local.centroids_update_interval > 1 << 60
        t = torch.tensor([float(rank + 1), 10.0])
        si._all_reduce(t)
        assert t.tolist() == [3.0, 20.0]
        g = si._all_gather(torch.tensor([[rank, rank * 7]]))
        assert g.shape == (2, 1, 2) and g[:, 0, 1].tolist() == [0, 7]
        # contiguous phase: ownership by range
        seeds = torch.tensor([0, 4, 5, 9, 12])
        mine, rows = si._local_rows_of(seeds)
        assert mine.tolist() == [lo <= s < hi for s in seeds.tolist()]
        assert rows.tolist() == [s - lo for s in seeds.tolist() if lo <= s < hi]
        # 7 online writes: ids 10..16 go to rank id % 2, appended to the owner's table in ascending order
        table = si._gid_table()
        new = [i for i in range(n0, n0 + 7) if i % world == rank]
        table[local.memory_count:local.memory_count + len(new)] = torch.tensor(new)
        local.memory_count += len(new)
        si.n_total = n0 + 7
        ids = torch.arange(0, 20)
        mine, rows = si._local_rows_of(ids)
        owned = [i for i in range(20) if (lo <= i < hi) or (n0 <= i < n0 + 7 and i % world == rank)]
        assert ids[mine].tolist() == owned
        assert table[rows].tolist() == owned
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_sharded_index_host_logic_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_index_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(q.get(timeout=5) for _ in range(2)) == [(0, "ok"), (1, "ok")]
