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
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None):
        self.rows = local_rows
        self.row_base = int(row_base)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if local_search is None or merge is None:
            from . import ops      # CUDA only; raises without the library
            if scale is None:
                scale = ops.row_inv_norms(local_rows)
            local_search = local_search or (lambda q, k: ops.scan_topk(self.rows, q, k, self.scale, self.bias,
                                                                       row_base=self.row_base))
            merge = merge or ops.topk_merge
        self.scale, self.bias = scale, bias
        self._local_search = local_search
        self._merge = merge

    def search(self, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries [B,d] identical on all ranks -> (global rows int64 [B,k], scores fp32 [B,k]) on all ranks."""
        if queries.device != self.rows.device:
            queries = queries.to(self.rows.device, non_blocking=True)
        idx, score = self._local_search(queries, k)
        if self.world == 1:
            return idx, score
        b = idx.shape[0]
        all_idx = torch.empty(self.world * b, k, dtype=idx.dtype, device=idx.device)
        all_score = torch.empty(self.world * b, k, dtype=score.dtype, device=score.device)
        dist.all_gather_into_tensor(all_idx, idx.contiguous(), group=self.group)      # rank-major concatenation
        dist.all_gather_into_tensor(all_score, score.contiguous(), group=self.group)
        all_idx, all_score = all_idx.view(self.world, b, k), all_score.view(self.world, b, k)
        # [G,B,k] -> [B,G*k] (the layout aura_topk_merge takes)
        cat_idx = all_idx.permute(1, 0, 2).reshape(b, self.world * k).contiguous()
        cat_score = all_score.permute(1, 0, 2).reshape(b, self.world * k).contiguous()
        out_score, out_idx = self._merge(cat_score, cat_idx, self.world, k, k)
        return out_idx, out_score
