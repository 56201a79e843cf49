"""Tensor-level wrappers over the C ABI: pointer/stream plumbing only, no arithmetic.

torch is used for device memory and streams; every computation happens inside
libaura_hippo.so.  All functions require CUDA tensors and raise otherwise.
"""
from __future__ import annotations

import functools
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import AURA_BF16, AURA_F32, AURA_IVF_EMPTY_OK, AURA_IVF_MEASURED_EPS, AURA_MAX_K, check


def _stream() -> int:
    """Handle of the current stream of the CURRENT device; every public entry runs under `_on_device`, which makes
    the tensors' device current first (the library launches on cudaGetDevice())."""
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument current: the library reads cudaGetDevice() for its
    launches and cached attributes, and `_stream()` / `_workspace` are per current device.  With the bank on cuda:1 and
    cuda:0 current, an unguarded call would launch on GPU 0 against GPU 1 pointers (the reference class works on any
    device index, hippocampal.py:50-53)."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapped


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return AURA_F32
    if t.dtype == torch.bfloat16:
        return AURA_BF16
    raise TypeError(f"memory bank rows must be float32 or bfloat16, got {t.dtype}")


def _dev(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.AuraLibraryError(f"{name} must be a CUDA tensor: the retrieval path has no CPU implementation")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_workspaces = {}
_pinned_streams = set()      # (device index, stream handle) whose scratch pointers are baked into a captured CUDA graph
_retired = []                # scratch tensors replaced on a pinned stream: kept alive, a graph may still replay into them


def pin_workspaces(device: torch.device, stream_handle: int) -> None:
    """Scratch buffers handed out on this stream are never freed from now on (`GraphedSearch`)."""
    _pinned_streams.add((device.index, int(stream_handle)))


def _workspace(nbytes: int, device: torch.device, tag: str = "ws") -> torch.Tensor:
    """Per-(device, stream) scratch, grown on demand (the library never allocates)."""
    st = _stream()
    key = (device.index, st, tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None and (device.index, st) in _pinned_streams:
            _retired.append(ws)
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


@_on_device
def row_inv_norms(rows: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows = _dev(rows, "rows")
    n, d = rows.shape
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=rows.device)
    check(_lib.load().aura_row_inv_norms(rows.data_ptr(), _dtype_code(rows), n, d, out.data_ptr(), _stream()),
          "aura_row_inv_norms")
    return out


@_on_device
def row_terms(metadata: torch.Tensor, inv_norm: torch.Tensor, now: float, n_rows: int,
              locations: Optional[torch.Tensor] = None, query_loc: Optional[torch.Tensor] = None,
              scale: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None):
    metadata = _dev(metadata, "metadata")
    dev = metadata.device
    if scale is None:
        scale = torch.empty(n_rows, dtype=torch.float32, device=dev)
    if bias is None:
        bias = torch.empty(n_rows, dtype=torch.float32, device=dev)
    sd = 0
    if query_loc is not None:
        locations = _dev(locations, "locations")
        query_loc = _dev(query_loc.to(device=dev, dtype=torch.float32), "query_loc")
        sd = locations.shape[1]
    check(_lib.load().aura_row_terms(metadata.data_ptr(), _ptr(locations) if query_loc is not None else None, sd,
                                     _ptr(query_loc), float(now), inv_norm.data_ptr(), n_rows, scale.data_ptr(),
                                     bias.data_ptr(), _stream()), "aura_row_terms")
    return scale, bias


@_on_device
def decay_strength(metadata: torch.Tensor, n_rows: int, rate: float) -> None:
    metadata = _dev(metadata, "metadata")
    check(_lib.load().aura_decay_strength(metadata.data_ptr(), n_rows, float(rate), _stream()), "aura_decay_strength")


@_on_device
def scan_topk(rows: torch.Tensor, queries: torch.Tensor, k: int, scale: Optional[torch.Tensor],
              bias: Optional[torch.Tensor] = None, n_rows: Optional[int] = None, row_base: int = 0,
              out_idx: Optional[torch.Tensor] = None, out_score: Optional[torch.Tensor] = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact scan + fused top-k.  queries [B,d] fp32 (raw; normalised inside).  Returns (idx[B,k], score[B,k])."""
    rows = _dev(rows, "rows")
    queries = _dev(queries, "queries")
    if queries.dtype != torch.float32:
        raise TypeError("queries must be float32")
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    b, d = queries.shape
    if rows.shape[1] != d:
        raise ValueError(f"query dim {d} != bank dim {rows.shape[1]}")
    n = rows.shape[0] if n_rows is None else int(n_rows)
    if not (1 <= k <= AURA_MAX_K):
        raise ValueError(f"k={k} outside [1,{AURA_MAX_K}]")
    dev = rows.device
    if out_idx is None:
        out_idx = torch.empty(b, k, dtype=torch.int64, device=dev)
    if out_score is None:
        out_score = torch.empty(b, k, dtype=torch.float32, device=dev)
    lib = _lib.load()
    nbytes = lib.aura_scan_topk_workspace_bytes(n, d, b, k)
    ws = _workspace(nbytes, dev)
    check(lib.aura_scan_topk(rows.data_ptr(), _dtype_code(rows), n, d, queries.data_ptr(), b, _ptr(scale), _ptr(bias),
                             k, row_base, out_idx.data_ptr(), out_score.data_ptr(), ws.data_ptr(), ws.numel(),
                             _stream()), "aura_scan_topk")
    return out_idx, out_score


@_on_device
def topk_merge(scores: torch.Tensor, idx: torch.Tensor, n_lists: int, k_in: int, k_out: int):
    scores = _dev(scores, "scores")
    idx = _dev(idx, "idx")
    b = scores.shape[0]
    out_s = torch.empty(b, k_out, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(b, k_out, dtype=torch.int64, device=scores.device)
    check(_lib.load().aura_topk_merge(scores.data_ptr(), idx.data_ptr(), b, n_lists, k_in, k_out, out_s.data_ptr(),
                                      out_i.data_ptr(), _stream()), "aura_topk_merge")
    return out_s, out_i


@_on_device
def pack_topk(idx: torch.Tensor, score: torch.Tensor, flags: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
              id_map: Optional[torch.Tensor] = None, id_base: int = 0) -> torch.Tensor:
    """[B, 2k+1] int64 payload of one rank for the sharded all-gather (ids, score bits, certification flag).
    Ids are id_map[idx] (int64 table: local row -> global memory id) or idx + id_base; missing results stay -1."""
    b, k = idx.shape
    if out is None:
        out = torch.empty(b, 2 * k + 1, dtype=torch.int64, device=idx.device)
    check(_lib.load().aura_pack_topk(idx.data_ptr(), score.data_ptr(), _ptr(flags), b, k, _ptr(id_map), int(id_base),
                                     out.data_ptr(), _stream()), "aura_pack_topk")
    return out


@_on_device
def topk_merge_packed(gathered: torch.Tensor, n_ranks: int, b: int, k: int):
    """Merge the rank-major gathered payloads [n_ranks*B, 2k+1] -> (idx [B,k], score [B,k], any_flag [B] int32)."""
    dev = gathered.device
    out_s = torch.empty(b, k, dtype=torch.float32, device=dev)
    out_i = torch.empty(b, k, dtype=torch.int64, device=dev)
    flag = torch.empty(b, dtype=torch.int32, device=dev)
    check(_lib.load().aura_topk_merge_packed(gathered.data_ptr(), n_ranks, b, k, out_s.data_ptr(), out_i.data_ptr(),
                                             flag.data_ptr(), _stream()), "aura_topk_merge_packed")
    return out_i, out_s, flag


def peer_gather_buffer_bytes(n_ranks: int, n_queries: int, k: int) -> int:
    return int(_lib.load().aura_peer_gather_buffer_bytes(n_ranks, n_queries, k))


@_on_device
def pack_scatter(idx: torch.Tensor, score: torch.Tensor, flags: Optional[torch.Tensor], peer_ptrs, rank: int, n_ranks: int,
                 counters: torch.Tensor, id_map: Optional[torch.Tensor] = None, id_base: int = 0) -> None:
    """Build this rank's [B, 2k+1] payload and store it into every rank's gather buffer (peer_ptrs: ctypes array of
    the peer-mapped buffer addresses in rank order), then raise this rank's flag there (aura_pack_scatter)."""
    b, k = idx.shape
    check(_lib.load().aura_pack_scatter(idx.data_ptr(), score.data_ptr(), _ptr(flags), b, k, _ptr(id_map), int(id_base),
                                        peer_ptrs, rank, n_ranks, counters.data_ptr(), _stream()), "aura_pack_scatter")


@_on_device
def merge_gathered(gather_buf: torch.Tensor, n_ranks: int, b: int, k: int, counters: torch.Tensor):
    """Wait (on the device) for every rank's payload of the current step in the local gather buffer, then merge:
    -> (idx [B,k], score [B,k], any_flag [B] int32)."""
    dev = gather_buf.device
    out_s = torch.empty(b, k, dtype=torch.float32, device=dev)
    out_i = torch.empty(b, k, dtype=torch.int64, device=dev)
    flag = torch.empty(b, dtype=torch.int32, device=dev)
    check(_lib.load().aura_merge_gathered(gather_buf.data_ptr(), n_ranks, b, k, counters.data_ptr(), out_s.data_ptr(),
                                          out_i.data_ptr(), flag.data_ptr(), _stream()), "aura_merge_gathered")
    return out_i, out_s, flag


@_on_device
def gather_rows(rows: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    rows = _dev(rows, "rows")
    idx = _dev(idx, "idx")
    d = rows.shape[1]
    out = torch.empty(*idx.shape, d, dtype=torch.float32, device=rows.device)
    check(_lib.load().aura_gather_rows(rows.data_ptr(), _dtype_code(rows), d, idx.data_ptr(), idx.numel(),
                                       out.data_ptr(), _stream()), "aura_gather_rows")
    return out


@_on_device
def gather_context(rows: torch.Tensor, idx: torch.Tensor, score: torch.Tensor, return_weights: bool = False):
    """context[b] = sum_j softmax(score[b])_j * rows[idx[b, j]] (missing results: score 0, zero row) - the memory context
    of inject_memories' "concat" / "gate" modes (memory_augmented_layer.py:185-194) fused with the gather."""
    rows = _dev(rows, "rows")
    idx = _dev(idx, "idx")
    score = _dev(score, "score")
    b, k = idx.shape
    d = rows.shape[1]
    ctx = torch.empty(b, d, dtype=torch.float32, device=rows.device)
    w = torch.empty(b, k, dtype=torch.float32, device=rows.device) if return_weights else None
    if k > 0:
        check(_lib.load().aura_gather_context(rows.data_ptr(), _dtype_code(rows), d, idx.data_ptr(), score.data_ptr(), b, k,
                                              ctx.data_ptr(), _ptr(w), _stream()), "aura_gather_context")
    else:
        ctx.zero_()
    return (ctx, w) if return_weights else ctx


# ---------------------------------------------------------------------------- bank write
@_on_device
def bank_write(rows: torch.Tensor, first_row: int, features: torch.Tensor, metadata: torch.Tensor,
               inv_norm: torch.Tensor, timestamp: float, locations: Optional[torch.Tensor] = None,
               location: Optional[torch.Tensor] = None) -> None:
    """Write features [n,d] (fp32, CUDA) into bank rows first_row.. plus metadata / location / inv_norm."""
    rows = _dev(rows, "rows")
    features = _dev(features, "features")
    if features.dtype != torch.float32:
        raise TypeError("features must be float32")
    if features.dim() == 1:
        features = features.unsqueeze(0)
    n, d = features.shape
    if d != rows.shape[1]:
        raise ValueError(f"feature dim {d} != bank dim {rows.shape[1]}")
    if first_row < 0 or first_row + n > rows.shape[0]:
        raise IndexError(f"rows [{first_row},{first_row + n}) outside the bank of {rows.shape[0]}")
    sd = 0 if locations is None else locations.shape[1]
    check(_lib.load().aura_bank_write(rows.data_ptr(), _dtype_code(rows), d, first_row, n, features.data_ptr(),
                                      _ptr(locations), sd, _ptr(location), metadata.data_ptr(), float(timestamp),
                                      inv_norm.data_ptr(), _stream()), "aura_bank_write")


# ---------------------------------------------------------------------------- centroid index
@_on_device
def kmeans_seed(rows: torch.Tensor, seed_rows: torch.Tensor, centroids: torch.Tensor) -> None:
    rows = _dev(rows, "rows")
    seed_rows = _dev(seed_rows, "seed_rows")
    if seed_rows.dtype != torch.int64:
        raise TypeError("seed_rows must be int64")
    check(_lib.load().aura_kmeans_seed(rows.data_ptr(), _dtype_code(rows), rows.shape[1], seed_rows.data_ptr(),
                                       seed_rows.numel(), centroids.data_ptr(), _stream()), "aura_kmeans_seed")


@_on_device
def kmeans_assign(rows: torch.Tensor, n_rows: int, centroids: torch.Tensor, n_centroids: int, assign: torch.Tensor,
                  cid_f32: Optional[torch.Tensor] = None, cid_stride: int = 1,
                  best: Optional[torch.Tensor] = None, inv_norm: Optional[torch.Tensor] = None) -> None:
    """assign[i] (int32) = nearest of centroids[:n_centroids]; optional float id written at cid_f32[i*cid_stride]."""
    rows = _dev(rows, "rows")
    lib = _lib.load()
    code = _dtype_code(rows)
    ws = _workspace(lib.aura_kmeans_assign_workspace_bytes(n_rows, rows.shape[1], code, n_centroids), rows.device, "assign")
    check(lib.aura_kmeans_assign(rows.data_ptr(), code, n_rows, rows.shape[1], centroids.data_ptr(), n_centroids,
                                 _ptr(inv_norm), assign.data_ptr(), _ptr(cid_f32), cid_stride, _ptr(best), ws.data_ptr(),
                                 ws.numel(), _stream()), "aura_kmeans_assign")


@_on_device
def ivf_build_lists(cid: torch.Tensor, n_rows: int, n_lists: int, list_offsets: torch.Tensor,
                    list_rows: torch.Tensor) -> None:
    cid = _dev(cid, "cid")
    lib = _lib.load()
    ws = _workspace(lib.aura_ivf_build_lists_workspace_bytes(n_lists), cid.device, "lists")
    check(lib.aura_ivf_build_lists(cid.data_ptr(), n_rows, n_lists, list_offsets.data_ptr(), list_rows.data_ptr(),
                                   ws.data_ptr(), ws.numel(), _stream()), "aura_ivf_build_lists")


@_on_device
def kmeans_list_sums(rows: torch.Tensor, list_offsets: torch.Tensor, list_rows: torch.Tensor, n_lists: int,
                     sums: torch.Tensor, counts: torch.Tensor) -> None:
    rows = _dev(rows, "rows")
    check(_lib.load().aura_kmeans_list_sums(rows.data_ptr(), _dtype_code(rows), rows.shape[1], list_offsets.data_ptr(),
                                            list_rows.data_ptr(), n_lists, sums.data_ptr(), counts.data_ptr(),
                                            _stream()), "aura_kmeans_list_sums")


@_on_device
def kmeans_finalize(sums: torch.Tensor, counts: torch.Tensor, n_centroids: int, centroids: torch.Tensor) -> None:
    check(_lib.load().aura_kmeans_finalize(sums.data_ptr(), counts.data_ptr(), n_centroids, centroids.shape[1],
                                           centroids.data_ptr(), _stream()), "aura_kmeans_finalize")


@_on_device
def ivf_list_counts(list_offsets: torch.Tensor, n_lists: int, counts_f32: torch.Tensor) -> None:
    check(_lib.load().aura_ivf_list_counts(list_offsets.data_ptr(), n_lists, counts_f32.data_ptr(), _stream()),
          "aura_ivf_list_counts")


@_on_device
def online_assign(rows: torch.Tensor, first_row: int, n_writes: int, centroids: torch.Tensor, n_live: int,
                  counts: torch.Tensor, cid_i32: torch.Tensor, cid_f32: Optional[torch.Tensor] = None,
                  cid_stride: int = 1) -> None:
    rows = _dev(rows, "rows")
    lib = _lib.load()
    ws = _workspace(lib.aura_online_assign_workspace_bytes(), rows.device, "online")
    check(lib.aura_online_assign(rows.data_ptr(), _dtype_code(rows), rows.shape[1], first_row, n_writes,
                                 centroids.data_ptr(), n_live, counts.data_ptr(), cid_i32.data_ptr(), _ptr(cid_f32),
                                 cid_stride, ws.data_ptr(), ws.numel(), _stream()), "aura_online_assign")


@_on_device
def ivf_coarse(queries: torch.Tensor, centroids: torch.Tensor, nprobe: int) -> torch.Tensor:
    queries = _dev(queries, "queries")
    centroids = _dev(centroids, "centroids")
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    b, d = queries.shape
    c = centroids.shape[0]
    probes = torch.empty(b, nprobe, dtype=torch.int64, device=queries.device)
    lib = _lib.load()
    ws = _workspace(lib.aura_ivf_coarse_workspace_bytes(b, d, c, nprobe), queries.device, "coarse")
    check(lib.aura_ivf_coarse(queries.data_ptr(), b, d, centroids.data_ptr(), c, nprobe, probes.data_ptr(),
                              ws.data_ptr(), ws.numel(), _stream()), "aura_ivf_coarse")
    return probes


@_on_device
def ivf_search(rows: torch.Tensor, n_rows: int, queries: torch.Tensor, centroids: torch.Tensor, nprobe: int,
               list_offsets: torch.Tensor, list_rows: torch.Tensor, k: int, scale: Optional[torch.Tensor],
               bias: Optional[torch.Tensor] = None, row_base: int = 0, return_probes: bool = False,
               allow_empty: bool = False):
    """Centroid-path query: coarse probes + scan of the probed lists.  Returns (idx[B,k], score[B,k][, probes]).
    allow_empty: a query whose probed lists hold no row returns idx -1 / score -inf instead of scanning every row
    (a row shard applies the reference's all-rows rule, hippocampal.py:269-270, to the MERGED result)."""
    rows = _dev(rows, "rows")
    queries = _dev(queries, "queries")
    if queries.dtype != torch.float32:
        raise TypeError("queries must be float32")
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    b, d = queries.shape
    if rows.shape[1] != d or centroids.shape[1] != d:
        raise ValueError("query / bank / centroid dims differ")
    if not (1 <= k <= AURA_MAX_K):
        raise ValueError(f"k={k} outside [1,{AURA_MAX_K}]")
    c = centroids.shape[0]
    dev = rows.device
    out_idx = torch.empty(b, k, dtype=torch.int64, device=dev)
    out_score = torch.empty(b, k, dtype=torch.float32, device=dev)
    probes = torch.empty(b, nprobe, dtype=torch.int64, device=dev) if return_probes else None
    lib = _lib.load()
    ws = _workspace(lib.aura_ivf_search_workspace_bytes(b, d, c, nprobe, k), dev)
    check(lib.aura_ivf_search(rows.data_ptr(), _dtype_code(rows), n_rows, d, queries.data_ptr(), b,
                              centroids.data_ptr(), c, nprobe, list_offsets.data_ptr(), list_rows.data_ptr(),
                              _ptr(scale), _ptr(bias), k, row_base, AURA_IVF_EMPTY_OK if allow_empty else 0,
                              out_idx.data_ptr(), out_score.data_ptr(), _ptr(probes), ws.data_ptr(), ws.numel(),
                              _stream()), "aura_ivf_search")
    return (out_idx, out_score, probes) if return_probes else (out_idx, out_score)


# ---------------------------------------------------------------------------- tensor-core paths
TC_MAX_K = 114                 # tensor-core paths: 32 candidates per round, up to 4 rounds, margin 14 (k <= 18: one round)
TC_MIN_BATCH = 3               # fp32 bank: below this the CUDA-core streaming scan is faster
TC_MIN_BATCH_BF16 = 2          # bf16 bank


def tc_min_batch(rows: torch.Tensor) -> int:
    """Smallest query block that goes to the tensor-core path.  B200, 1M x 768: the scan takes 470 / 498 / 687 us at
    B = 1 / 2 / 3 (fp32) and 273 / 376 / 748 us (bf16); the tensor-core path 530-560 us (fp32) and 315-366 us (bf16) for
    any B <= 128."""
    return TC_MIN_BATCH_BF16 if rows.dtype == torch.bfloat16 else TC_MIN_BATCH
TC_EPS_COS = 2.0 ** -9 + 1e-4   # |tensor-core cosine - fp32 cosine| bound: both operands rounded to 11 bits + fp32 sums
TC_EPS_COS_BF16 = 2.0 ** -7 + 1e-4   # the same bound when both operands are rounded to bf16 (8 significant bits)
TC_SHADOW_MAX_K = 34            # bf16 shadow shortlist keeps 48 candidates per query (k + 14 margin)


def batch_topk_supported(rows: torch.Tensor, k: int) -> bool:
    return (rows.shape[1] * rows.element_size()) % 16 == 0 and rows.data_ptr() % 16 == 0 and 1 <= k <= TC_MAX_K


class Bf16Shadow:
    """bf16 copy of an fp32 bank for the tensor-core shortlist pass + the measured bound on its rounding error
    (`relerr`, device scalar: max over converted rows of ||bf16(r) - r|| / ||r||; only ever raised)."""

    def __init__(self, rows: torch.Tensor, n_rows: Optional[int] = None):
        self.rows = torch.empty(rows.shape, dtype=torch.bfloat16, device=rows.device)
        self.relerr = torch.zeros(1, dtype=torch.float32, device=rows.device)
        self.refresh(rows, 0, rows.shape[0] if n_rows is None else n_rows)

    def refresh(self, rows: torch.Tensor, lo: int, hi: int) -> None:
        if hi > lo:
            rows_to_bf16(rows, self.rows, first_row=lo, n_rows=hi - lo, relerr=self.relerr)


@_on_device
def rows_to_bf16(rows: torch.Tensor, out: Optional[torch.Tensor] = None, first_row: int = 0,
                 n_rows: Optional[int] = None, relerr: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 shadow copy of fp32 bank rows [first_row, first_row + n_rows) (whole tensor by default)."""
    rows = _dev(rows, "rows")
    if rows.dtype != torch.float32:
        raise TypeError("rows must be float32")
    n = rows.shape[0] - first_row if n_rows is None else int(n_rows)
    if out is None:
        out = torch.empty(rows.shape, dtype=torch.bfloat16, device=rows.device)
    if n > 0:
        check(_lib.load().aura_rows_to_bf16(rows[first_row:].data_ptr(), n, rows.shape[1], out[first_row:].data_ptr(),
                                            _ptr(relerr), _stream()), "aura_rows_to_bf16")
    return out


@_on_device
def batch_topk(rows: torch.Tensor, queries: torch.Tensor, k: int, scale: Optional[torch.Tensor],
               bias: Optional[torch.Tensor] = None, n_rows: Optional[int] = None, row_base: int = 0,
               eps: float = TC_EPS_COS, shadow=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Tensor-core shortlist + exact fp32 re-score.  Returns (idx[B,k], score[B,k], uncertain[B] int32).
    shadow: a `Bf16Shadow` of the fp32 bank (or a bare bf16 tensor); the shortlist pass then reads it instead of the fp32
    rows.  With a `Bf16Shadow`, `eps` is the score-per-cosine unit (max |scale_r| * ||r||: 1 for pure cosine) and the
    certification bound is measured per query from the shadow's rounding error; with a bare tensor `eps` must be the
    worst-case bound (TC_EPS_COS_BF16 per unit)."""
    rows = _dev(rows, "rows")
    queries = _dev(queries, "queries")
    if queries.dtype != torch.float32:
        raise TypeError("queries must be float32")
    b, d = queries.shape
    n = rows.shape[0] if n_rows is None else int(n_rows)
    dev = rows.device
    out_idx = torch.empty(b, k, dtype=torch.int64, device=dev)
    out_score = torch.empty(b, k, dtype=torch.float32, device=dev)
    flags = torch.empty(b, dtype=torch.int32, device=dev)
    lib = _lib.load()
    code = _dtype_code(rows)
    nbytes = lib.aura_batch_topk_workspace_bytes(n, d, code, b, k)
    if nbytes == 0:
        raise _lib.AuraLibraryError(f"aura_batch_topk: unsupported shape n={n} d={d} B={b} k={k}")
    ws = _workspace(nbytes, dev, "batch")
    relerr = None
    if isinstance(shadow, Bf16Shadow):
        shadow, relerr = shadow.rows, shadow.relerr
    if shadow is not None and (shadow.dtype != torch.bfloat16 or shadow.shape[1] != d or shadow.shape[0] < n
                               or not shadow.is_contiguous() or rows.dtype != torch.float32 or k > TC_SHADOW_MAX_K):
        raise ValueError("bf16 shadow does not match the bank (fp32 rows, same shape, k <= TC_SHADOW_MAX_K)")
    check(lib.aura_batch_topk(rows.data_ptr(), code, n, d, queries.data_ptr(), b, _ptr(scale), _ptr(bias), k, row_base,
                              float(eps), _ptr(shadow), _ptr(relerr), out_idx.data_ptr(), out_score.data_ptr(),
                              flags.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "aura_batch_topk")
    return out_idx, out_score, flags


@_on_device
def exact_topk_batched(rows: torch.Tensor, queries: torch.Tensor, k: int, scale: Optional[torch.Tensor],
                       bias: Optional[torch.Tensor] = None, n_rows: Optional[int] = None, row_base: int = 0,
                       eps: float = TC_EPS_COS, stats: Optional[dict] = None, defer: bool = False, shadow=None):
    """Exact top-k of a query block: tensor-core pass, then the exact streaming scan for the (rare) queries
    whose result could not be certified.  One small D2H read (the flag count) per call.

    defer=True returns (idx, score, flags) WITHOUT reading the flags: the caller must later call
    `exact_topk_fixup` with the same arguments for the flagged queries (ShardedBank does this after it has
    enqueued its collectives, so that the only host sync of a step comes after all of its work is queued)."""
    idx, score, flags = batch_topk(rows, queries, k, scale, bias, n_rows, row_base, eps, shadow)
    if defer:
        return idx, score, flags
    exact_topk_fixup(flags, idx, score, rows, queries, k, scale, bias, n_rows, row_base, stats)
    return idx, score


@_on_device
def exact_topk_fixup(flags: torch.Tensor, idx: torch.Tensor, score: torch.Tensor, rows: torch.Tensor,
                     queries: torch.Tensor, k: int, scale: Optional[torch.Tensor], bias: Optional[torch.Tensor] = None,
                     n_rows: Optional[int] = None, row_base: int = 0, stats: Optional[dict] = None) -> int:
    """Re-run the flagged queries through the exact scan, in place.  Returns how many were flagged (host sync)."""
    bad = torch.nonzero(flags, as_tuple=False).squeeze(-1)
    if stats is not None:
        stats["uncertain"] = stats.get("uncertain", 0) + int(bad.numel())
    if bad.numel() > 0:
        i2, s2 = scan_topk(rows, queries[bad].contiguous(), k, scale, bias, n_rows=n_rows, row_base=row_base)
        idx[bad] = i2
        score[bad] = s2
    return int(bad.numel())


@_on_device
def allpairs_topk(rows: torch.Tensor, k: int = 32, inv_norm: Optional[torch.Tensor] = None,
                  n_rows: Optional[int] = None, a_first: int = 0, n_a: Optional[int] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Cognitive map: top-k cosine neighbours (self excluded) of rows[a_first:a_first+n_a] within rows[:n_rows]."""
    rows = _dev(rows, "rows")
    n = rows.shape[0] if n_rows is None else int(n_rows)
    na = n - a_first if n_a is None else int(n_a)
    d = rows.shape[1]
    dev = rows.device
    if inv_norm is None:
        inv_norm = row_inv_norms(rows)
    out_idx = torch.empty(na, k, dtype=torch.int64, device=dev)
    out_score = torch.empty(na, k, dtype=torch.float32, device=dev)
    lib = _lib.load()
    code = _dtype_code(rows)
    nbytes = lib.aura_allpairs_topk_workspace_bytes(na, n, d, code, k)
    if nbytes == 0:
        raise _lib.AuraLibraryError(f"aura_allpairs_topk: unsupported shape n={n} d={d} k={k}")
    ws = _workspace(nbytes, dev, "allpairs")
    check(lib.aura_allpairs_topk(rows.data_ptr(), code, n, d, a_first, na, inv_norm.data_ptr(), k, out_idx.data_ptr(),
                                 out_score.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "aura_allpairs_topk")
    return out_idx, out_score


@_on_device
def ivf_pack_lists(rows: torch.Tensor, list_rows: torch.Tensor, n_listed: int, out: torch.Tensor,
                   relerr: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[p] = rows[list_rows[p]], p < n_listed: the list-major resident copy of the bank (aura_ivf_pack_lists).
    `out` has the bank's dtype, or is bfloat16 over an fp32 bank (a list-major bf16 SHADOW: half the bytes for the
    tensor-core pass); `relerr` (device float[1]) is then reset and raised to the largest relative rounding error
    ||bf16(r) - r|| / ||r|| of the packed rows, which the search turns into its per-query certification bound."""
    rows = _dev(rows, "rows")
    shadow = out.dtype == torch.bfloat16 and rows.dtype == torch.float32
    if (out.dtype != rows.dtype and not shadow) or out.shape[1] != rows.shape[1] or out.shape[0] < n_listed or not out.is_contiguous():
        raise ValueError("out must be a contiguous [>= n_listed, d] tensor of the bank's dtype (or bfloat16 over an fp32 bank)")
    if shadow:
        if relerr is None or relerr.dtype != torch.float32 or relerr.numel() != 1 or relerr.device != rows.device:
            raise ValueError("a bf16 list-major shadow needs relerr: a float32 device tensor of one element")
        relerr.zero_()
    check(_lib.load().aura_ivf_pack_lists(rows.data_ptr(), _dtype_code(rows), rows.shape[1], list_rows.data_ptr(), n_listed,
                                          out.data_ptr(), _dtype_code(out), _ptr(relerr) if shadow else None, _stream()),
          "aura_ivf_pack_lists")
    return out


TC_IVF_MIN_BATCH = 64          # list-major grouped GEMM pays off once lists are shared by several queries


@_on_device
def ivf_search_batched(rows: torch.Tensor, n_rows: int, queries: torch.Tensor, centroids: torch.Tensor, nprobe: int,
                       list_offsets: torch.Tensor, list_rows: torch.Tensor, k: int, scale: Optional[torch.Tensor],
                       bias: Optional[torch.Tensor] = None, row_base: int = 0, eps: float = TC_EPS_COS,
                       stats: Optional[dict] = None, strict: bool = True,
                       rows_by_list: Optional[torch.Tensor] = None, allow_empty: bool = False,
                       lm_relerr: Optional[torch.Tensor] = None, measured_eps: bool = False
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Centroid-path query for a block of queries: list-major tensor-core pass (aura_ivf_search_batch), then the
    per-query path for the queries it hands back.

    rows_by_list : optional resident copy of the bank in list order (`ivf_pack_lists`); list tiles are then streamed
                   by TMA instead of gathered row by row.  Results do not depend on it.  A bfloat16 copy of an fp32
                   bank (with `lm_relerr`, both from `ivf_pack_lists`) halves the list bytes and doubles the tensor
                   rate; `eps` is then the score-per-cosine unit (max |scale_r| * ||r||) and the certification bound
                   is measured per query, as in `batch_topk` with a `Bf16Shadow`.  Results are re-scored from the
                   fp32 rows either way.
    measured_eps : bf16 bank: `eps` is the score-per-cosine unit and the bound is measured per query from the rounding
                   error of the bf16 query copy (the rows are exact operands); ~1.7x tighter than TC_EPS_COS.

    strict=True  : every flagged query (result not certified exact among its candidates, or no candidates) is re-run
                   through the per-query path - results equal `ivf_search` bit for bit.
    strict=False : only queries WITHOUT candidates are re-run (the reference then scans all rows, hippocampal.py:
                   269-270); uncertified ones keep the exact-re-scored tensor-core shortlist.  The scores returned are
                   still exact fp32; a candidate can only be missed if its TF32/bf16 score fell more than the shortlist
                   margin (>= 14 places) below its exact rank - near-tie heavy data (thousands of near-duplicates per
                   query) is where this mode saves the per-query re-runs."""
    rows = _dev(rows, "rows")
    queries = _dev(queries, "queries")
    if queries.dtype != torch.float32:
        raise TypeError("queries must be float32")
    b, d = queries.shape
    c = centroids.shape[0]
    dev = rows.device
    out_idx = torch.empty(b, k, dtype=torch.int64, device=dev)
    out_score = torch.empty(b, k, dtype=torch.float32, device=dev)
    flags = torch.empty(b, dtype=torch.int32, device=dev)
    lib = _lib.load()
    ws = _workspace(lib.aura_ivf_search_batch_workspace_bytes(b, d, c, nprobe, k), dev, "ivfbatch")
    lm_code = _dtype_code(rows_by_list) if rows_by_list is not None else _dtype_code(rows)
    lm_shadow = rows_by_list is not None and rows_by_list.dtype == torch.bfloat16 and rows.dtype == torch.float32
    if lm_shadow and lm_relerr is None:
        raise ValueError("a bf16 list-major shadow of an fp32 bank needs lm_relerr (ivf_pack_lists)")
    check(lib.aura_ivf_search_batch(rows.data_ptr(), _dtype_code(rows), n_rows, d, queries.data_ptr(), b,
                                    centroids.data_ptr(), c, nprobe, list_offsets.data_ptr(), list_rows.data_ptr(),
                                    _ptr(rows_by_list), lm_code, _ptr(lm_relerr) if lm_shadow else None,
                                    _ptr(scale), _ptr(bias), k, row_base,
                                    (AURA_IVF_EMPTY_OK if allow_empty else 0) |
                                    (AURA_IVF_MEASURED_EPS if measured_eps and rows.dtype == torch.bfloat16 else 0),
                                    float(eps), out_idx.data_ptr(),
                                    out_score.data_ptr(), flags.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
          "aura_ivf_search_batch")
    if stats is not None:
        import ctypes as _C
        import os as _os
        forced = _os.environ.get("AURA_IVF_ROWS")
        rows_path = ((forced != "0") if forced is not None else (k + 14 > 32)) or ((lm_shadow or (measured_eps and rows.dtype == torch.bfloat16)) and k + 14 > 32)   # the library's dispatch rule (ivf_batch.cu)
        stats["path"] = "rows-as-M, one-pass selection" if rows_path else "queries-as-M, register lists"
        stats["handed_back"] = int(flags.sum())
        if rows_path:                       # work-table / result-slot diagnostics exist for this formulation only
            items, cap = torch.zeros(8, dtype=torch.int32, device=dev), _C.c_int32(0)
            check(lib.aura_ivf_search_batch_items(ws.data_ptr(), b, d, c, nprobe, k, items.data_ptr(), _C.addressof(cap), _stream()),
                  "aura_ivf_search_batch_items")
            it = items.tolist()
            stats["items"], stats["items_cap"] = it[0], cap.value
            stats["slots"], stats["slots_per_query_max"], stats["kernel_flagged"], stats["no_candidates"] = it[1], it[2], it[3], it[4]
    if not strict:
        # keep only "no candidate at all" (never flagged when allow_empty) and work-table overflow
        flags = flags * (out_idx[:, 0] < 0).to(flags.dtype)
    bad = torch.nonzero(flags, as_tuple=False).squeeze(-1)
    if stats is not None:
        stats["uncertain"] = stats.get("uncertain", 0) + int(bad.numel())
    if bad.numel() > 0:
        i2, s2 = ivf_search(rows, n_rows, queries[bad].contiguous(), centroids, nprobe, list_offsets, list_rows, k, scale,
                            bias, row_base, allow_empty=allow_empty)
        out_idx[bad] = i2
        out_score[bad] = s2
    return out_idx, out_score
