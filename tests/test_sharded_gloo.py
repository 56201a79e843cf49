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
