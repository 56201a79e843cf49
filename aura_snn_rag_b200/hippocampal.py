"""Drop-in `HippocampalFormation` whose memory bank + centroid index run in libaura_hippo.so.

Mirrors the public surface of the reference class (src/core/hippocampal.py:31-377): constructor
signature (:41-49), the 13 registered buffers with the same names / shapes (state-dict
compatible), `create_episodic_memory` (:195-243), `retrieve_similar_memories` (:245-319),
`decay_memories` / `decay` (:321-343), `rebuild_centroids` (:345-377), the spatial / temporal
context helpers (:120-193) and the attributes callers touch directly (`memory_count`,
`id_to_idx`, `episodic_memories`, `centroids_k`, `centroids_update_interval`, `_index_ready`).

What is different underneath
* every O(M) or O(M*d) step is one hand-written sm_100a kernel behind the C ABI
  (include/aura_hippo.h); torch supplies device memory and streams only;
* there is NO CPU path: constructing the module without CUDA raises (the reference silently falls
  back to CPU, :53);
* inverted lists (CSR) replace the per-query mask passes (:264-268); per-row inverse norms and
  score terms replace the per-query re-normalisation of the whole bank (:278);
* result rows of the centroid path are bank rows (the reference maps candidate-local positions
  through id_to_idx, :307-317 - a bug); `k` is clamped to the candidate
  count; `location=` works together with the centroid path.  Scores are the reference's.
* batched / tensor-valued entry points exist next to the reference API: `retrieve_batch`,
  `exact_topk`, `create_episodic_memories`.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import AURA_MAX_K, AuraLibraryError
from .idtable import IdTable


@dataclass
class EpisodicMemory:
    """Per-memory metadata record, as hippocampal.py:23-29."""
    memory_id: str
    feature_idx: int
    timestamp: float
    strength: float = 1.0


class HippocampalFormation(nn.Module):
    def __init__(self,
                 spatial_dimensions: int = 2,
                 n_place_cells: int = 2000,
                 n_time_cells: int = 100,
                 n_grid_cells: int = 200,
                 max_memories: int = 100000,
                 feature_dim: int = 768,
                 device: str = 'cuda',
                 use_centroid_index: bool = True,
                 *,
                 centroids_k: int = 256,
                 centroid_rows: Optional[int] = None,
                 nprobe: int = 8,
                 bank_dtype: torch.dtype = torch.float32,
                 track_ids: bool = True,
                 list_major_copy: Union[bool, str] = False,
                 bf16_shadow: bool = False):
        super().__init__()
        if not torch.cuda.is_available():
            raise AuraLibraryError("HippocampalFormation (B200 build) needs a CUDA device: the retrieval path has "
                                   "no CPU implementation")
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise AuraLibraryError(f"device={device!r}: the retrieval path has no CPU implementation")
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        if bank_dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("bank_dtype must be torch.float32 or torch.bfloat16")
        self.spatial_dims = spatial_dimensions
        self.device = dev
        f32 = dict(device=dev, dtype=torch.float32)

        # place / grid / time cells (hippocampal.py:55-82): small state, same buffer names
        self.register_buffer('place_centers', torch.rand(n_place_cells, spatial_dimensions, **f32) * 20 - 10)
        self.register_buffer('place_radii', torch.rand(n_place_cells, 1, **f32) * 1.5 + 0.5)
        self.place_max_rate = 20.0
        spacings = torch.logspace(0, 2, n_grid_cells, base=2.0, **f32).unsqueeze(1)
        self.register_buffer('grid_spacings', spacings)
        self.register_buffer('grid_orientations', torch.rand(n_grid_cells, 1, **f32) * (torch.pi / 3))
        self.register_buffer('grid_phases', torch.rand(n_grid_cells, spatial_dimensions, **f32) * spacings)
        self.grid_max_rate = 25.0
        intervals = torch.logspace(0, 3, n_time_cells, base=10.0, **f32).unsqueeze(1)
        self.register_buffer('time_intervals', intervals)
        self.register_buffer('time_widths', intervals * 0.3)

        # memory bank (hippocampal.py:84-99)
        self.max_memories = int(max_memories)
        self.memory_count = 0
        self.register_buffer('memory_features', torch.zeros(max_memories, feature_dim, device=dev, dtype=bank_dtype))
        self.register_buffer('memory_locations', torch.zeros(max_memories, spatial_dimensions, **f32))
        self.register_buffer('memory_metadata', torch.zeros(max_memories, 4, **f32))
        self.episodic_memories: Dict[str, EpisodicMemory] = {}
        self.track_ids = bool(track_ids)
        self._ids = IdTable()
        self.current_location = torch.zeros(spatial_dimensions, **f32)
        self.last_event_time = time.time()
        self.register_buffer('k_const', 4 * torch.pi / torch.sqrt(torch.tensor(3.0, **f32)))

        # centroid index (hippocampal.py:112-118); knobs stay plain mutable attributes
        self.use_centroid_index = use_centroid_index
        self.centroids_k = int(centroids_k)
        self.centroids_update_interval = 512
        self.nprobe = int(nprobe)                       # reference literal 8 (:262)
        # batched centroid path: keep a second copy of the bank with every inverted list contiguous (2x bank memory), so
        # list tiles stream from HBM by TMA instead of being gathered row by row; re-packed after every list rebuild
        # list_major_copy="bf16" over an fp32 bank keeps that copy in bf16 (a list-major SHADOW, +0.5x bank memory): the
        # tensor-core pass reads half the bytes at twice the rate, results are still re-scored from the fp32 rows and
        # certified with a per-query bound measured from the copy's rounding error
        self._bank_by_list: Optional[torch.Tensor] = None
        self._lm_relerr: Optional[torch.Tensor] = None
        self._by_list_valid = False
        self.set_list_major_copy(list_major_copy)
        self.ivf_strict = True                          # batched centroid path: re-run uncertified queries exactly (ops.ivf_search_batched)
        # exact search of query blocks over an fp32 bank: shortlist on a bf16 copy of the bank (+50 % memory, twice the
        # tensor rate, half the bytes), exact fp32 re-score from the fp32 rows - same certified results
        self.bf16_shadow = bool(bf16_shadow) and bank_dtype == torch.float32
        self._shadow: Optional[ops.Bf16Shadow] = None
        self._shadow_dirty = [0, 0]                     # [lo, hi) rows written since the shadow was last refreshed
        rows = int(centroid_rows) if centroid_rows is not None else max(256, self.centroids_k)
        self.register_buffer('centroids', torch.zeros(rows, feature_dim, **f32))
        self.register_buffer('centroid_counts', torch.zeros(rows, **f32))
        self._index_ready = False

        # derived index state (not part of the reference's state dict)
        self.register_buffer('_inv_norm', torch.zeros(max_memories, **f32), persistent=False)
        self.register_buffer('_cid', torch.full((max_memories,), -1, device=dev, dtype=torch.int32), persistent=False)
        self.register_buffer('_list_offsets', torch.zeros(rows + 1, device=dev, dtype=torch.int32), persistent=False)
        self.register_buffer('_list_rows', torch.zeros(max_memories, device=dev, dtype=torch.int32), persistent=False)
        self.register_buffer('_scale', torch.zeros(max_memories, **f32), persistent=False)
        self.register_buffer('_bias', torch.zeros(max_memories, **f32), persistent=False)
        self._lists_dirty = True
        self._host_stage: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}   # pinned result staging per k
        self._terms_key = None        # (fp32 now, location key, version) the cached _scale/_bias belong to
        self._version = 0             # bumped by every write / decay
        # upper bound of the live strengths, tracked on the host (a write sets strength 1, :215; decay scales all of
        # them, :334): the tensor-core error bound scales with it and must not cost a device read per write
        self._strength_bound = 1.0
        self._derived_stale = False   # set by load_state_dict: derived buffers are recomputed before the next use

    # ------------------------------------------------------------------ reference attribute surface
    @property
    def id_to_idx(self) -> Dict[str, int]:
        return self._ids.id_to_idx

    # ------------------------------------------------------------------ spatial / temporal context
    def _to_dev_f32(self, x) -> torch.Tensor:
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        return x.to(device=self.device, dtype=torch.float32, non_blocking=True)

    def update_spatial_state(self, new_location, dt: float = 0.1) -> None:
        """hippocampal.py:120-132."""
        loc = self._to_dev_f32(new_location)
        self.current_location = loc if loc.dim() == 1 else loc[0]

    def get_spatial_context(self) -> Dict[str, Any]:
        """Place / grid cell rates at the current location (hippocampal.py:134-179); not on the hot path."""
        loc = self.current_location.unsqueeze(0)
        dist = torch.norm(loc - self.place_centers, dim=1, keepdim=True)
        sigma = self.place_radii / 3.0
        place = self.place_max_rate * torch.exp(-(dist ** 2) / (2 * sigma ** 2)) * (dist <= self.place_radii).float()
        c, s = torch.cos(self.grid_orientations), torch.sin(self.grid_orientations)
        x, y = loc[0, 0], loc[0, 1]
        rot = torch.cat([c * x - s * y, s * x + c * y], dim=1) - self.grid_phases
        kk = self.k_const / self.grid_spacings
        a, b = rot[:, 0:1], rot[:, 1:2]
        waves = torch.cos(kk * a) + torch.cos(kk * (-0.5 * a + 0.866 * b)) + torch.cos(kk * (-0.5 * a - 0.866 * b))
        grid = self.grid_max_rate * torch.relu(waves / 3.0 + 0.5)
        return {"current_location": self.current_location, "place_cells": place.flatten(),
                "grid_cells": grid.flatten(), "n_memories": self.memory_count}

    def get_temporal_context(self) -> Dict[str, Any]:
        """Time-cell rates (hippocampal.py:181-193)."""
        elapsed = time.time() - self.last_event_time
        diff = elapsed - self.time_intervals
        rates = 15.0 * torch.exp(-(diff ** 2) / (2 * (self.time_widths / 3) ** 2))
        return {"time_cells": rates.flatten(), "elapsed": elapsed}

    # ------------------------------------------------------------------ writes
    def _centroid_buffer_rows(self) -> int:
        return self.centroids.shape[0]

    def create_episodic_memory(self, memory_id: str, event_id: str, features,
                               associated_experts: Optional[List[str]] = None, *,
                               seed_rows: Optional[torch.Tensor] = None) -> None:
        """One-shot write + online k-means step (hippocampal.py:195-243).

        `seed_rows` (extension) is handed to a rebuild this write triggers, so that tests can replay the
        reference's randperm draw; None draws torch.randperm on the device like the reference.
        """
        if self.memory_count >= self.max_memories:
            idx = self.memory_count % self.max_memories     # full-bank quirk kept: always row 0 (:200-202)
        else:
            idx = self.memory_count
            self.memory_count += 1
        self._ensure_derived()
        self._strength_bound = max(self._strength_bound, 1.0)
        feats = self._to_dev_f32(features).detach().reshape(1, -1).contiguous()
        now = time.time()
        ops.bank_write(self.memory_features, idx, feats, self.memory_metadata, self._inv_norm, now,
                       self.memory_locations, self.current_location.contiguous())
        self._mark_shadow_dirty(idx, idx + 1)
        if self.use_centroid_index and self._index_ready:
            live = min(self.centroids_k, self._centroid_buffer_rows())          # :220
            ops.online_assign(self.memory_features, idx, 1, self.centroids, live, self.centroid_counts, self._cid,
                              self.memory_metadata[:, 2], 4)
        else:
            self._cid[idx:idx + 1].fill_(-1)
        self._lists_dirty = True
        self._version += 1
        if self.track_ids:
            self.episodic_memories[memory_id] = EpisodicMemory(memory_id=memory_id, feature_idx=idx, timestamp=now)
            self._ids.set(memory_id, idx)
        if (self.use_centroid_index and self.memory_count % self.centroids_update_interval == 0
                and self.memory_count > self.centroids_k):              # :242-243
            self.rebuild_centroids(seed_rows=seed_rows)

    def create_episodic_memories(self, features, memory_ids: Optional[Sequence[str]] = None) -> Tuple[int, int]:
        """Bulk append of N rows: same semantics as N calls of
        `create_episodic_memory` with one timestamp, including the sequential online-assign order and the
        periodic rebuild points.  Returns the [first, last+1) row range written.  The bank must have room
        (no FIFO overwrite in bulk mode)."""
        self._ensure_derived()
        self._strength_bound = max(self._strength_bound, 1.0)
        feats = self._to_dev_f32(features).detach()
        if feats.dim() == 1:
            feats = feats.unsqueeze(0)
        feats = feats.contiguous()
        n = feats.shape[0]
        first = self.memory_count
        if first + n > self.max_memories:
            raise ValueError(f"bulk write of {n} rows does not fit: {first} of {self.max_memories} rows used")
        if memory_ids is not None and len(memory_ids) != n:
            raise ValueError("memory_ids length differs from the number of rows")
        now = time.time()
        ops.bank_write(self.memory_features, first, feats, self.memory_metadata, self._inv_norm, now,
                       self.memory_locations, self.current_location.contiguous())
        self._mark_shadow_dirty(first, first + n)
        self._cid[first:first + n].fill_(-1)
        done = 0
        interval = self.centroids_update_interval
        while done < n:
            # rows up to the next rebuild point are assigned online (if the index is ready), then rebuild
            count = self.memory_count
            nxt = (count // interval + 1) * interval
            step = min(n - done, nxt - count)
            if self.use_centroid_index and self._index_ready:
                live = min(self.centroids_k, self._centroid_buffer_rows())
                ops.online_assign(self.memory_features, first + done, step, self.centroids, live,
                                  self.centroid_counts, self._cid, self.memory_metadata[:, 2], 4)
            self.memory_count += step
            done += step
            if (self.use_centroid_index and self.memory_count % interval == 0
                    and self.memory_count > self.centroids_k):
                self.rebuild_centroids()
        self._lists_dirty = True
        self._version += 1
        if self.track_ids and memory_ids is not None:
            for j, mid in enumerate(memory_ids):
                self.episodic_memories[mid] = EpisodicMemory(memory_id=mid, feature_idx=first + j, timestamp=now)
                self._ids.set(mid, first + j)
        return first, first + n

    def decay_memories(self, decay_rate: float = 0.01) -> None:
        """strength *= (1 - rate) over live rows (hippocampal.py:321-334)."""
        if self.memory_count == 0:
            return
        ops.decay_strength(self.memory_metadata, self.memory_count, decay_rate)
        self._strength_bound *= abs(1.0 - float(decay_rate))
        self._version += 1

    def decay(self, rate: float = 0.01) -> None:
        self.decay_memories(decay_rate=rate)

    # ------------------------------------------------------------------ index build
    def rebuild_centroids(self, seed_rows: Optional[torch.Tensor] = None) -> None:
        """Sample k seeds, one Lloyd step, re-assign, counts (hippocampal.py:345-377)."""
        if self.memory_count == 0 or not self.use_centroid_index:
            return
        self._ensure_derived()
        m = self.memory_count
        k = min(self.centroids_k, m)
        rows_c = self._centroid_buffer_rows()
        if k > rows_c:
            raise ValueError(f"centroids_k={self.centroids_k} exceeds the centroid buffer ({rows_c} rows); "
                             "pass centroids_k / centroid_rows to the constructor")
        if seed_rows is None:
            seed_rows = torch.randperm(m, device=self.device)[:k]          # :354
        seeds = seed_rows.to(device=self.device, dtype=torch.int64)[:k].contiguous()
        bank = self.memory_features
        ops.kmeans_seed(bank, seeds, self.centroids)                          # :355
        ops.kmeans_assign(bank, m, self.centroids, k, self._cid, inv_norm=self._inv_norm)   # :358-359
        ops.ivf_build_lists(self._cid, m, rows_c, self._list_offsets, self._list_rows)
        sums = torch.empty(rows_c, bank.shape[1], device=self.device, dtype=torch.float64)
        counts = torch.empty(rows_c, device=self.device, dtype=torch.int64)
        ops.kmeans_list_sums(bank, self._list_offsets, self._list_rows, rows_c, sums, counts)
        ops.kmeans_finalize(sums, counts, k, self.centroids)                  # :360-363 (empty keeps its seed)
        if k < self.centroids_k and k < rows_c:
            self.centroids[k:].zero_()                                        # :366-367
        ops.kmeans_assign(bank, m, self.centroids, k, self._cid, self.memory_metadata[:, 2], 4,
                          inv_norm=self._inv_norm)                              # :370-371,:376
        ops.ivf_build_lists(self._cid, m, rows_c, self._list_offsets, self._list_rows)
        counts_f = torch.zeros(max(self.centroids_k, 1), device=self.device, dtype=torch.float32)
        full = torch.empty(rows_c, device=self.device, dtype=torch.float32)
        ops.ivf_list_counts(self._list_offsets, rows_c, full)
        counts_f[:min(rows_c, counts_f.numel())] = full[:counts_f.numel()]
        self.centroid_counts = counts_f                                       # rebinding quirk kept (:369,:374)
        self._lists_dirty = False
        self._by_list_valid = False
        self._index_ready = True

    def _ensure_lists(self) -> None:
        if self._lists_dirty:
            ops.ivf_build_lists(self._cid, self.memory_count, self._centroid_buffer_rows(), self._list_offsets,
                                self._list_rows)
            self._lists_dirty = False
            self._by_list_valid = False

    def set_list_major_copy(self, mode: Union[bool, str]) -> None:
        """False: no copy; True / 'bank': a copy in the bank's dtype; 'bf16': a bf16 copy (a shadow when the bank is fp32).
        Drops the copy held for another mode; the next batched centroid-path query packs the new one."""
        if isinstance(mode, str) and mode.lower() not in ("bf16", "bank"):
            raise ValueError("list_major_copy must be a bool, 'bank' or 'bf16'")
        bank = self.memory_features
        lm_bf16 = (isinstance(mode, str) and mode.lower() == "bf16" and bank.dtype == torch.float32
                   and bank.shape[1] % 8 == 0)
        if lm_bf16 != getattr(self, "_lm_bf16", None) or not mode:
            self._bank_by_list, self._lm_relerr, self._by_list_valid = None, None, False
        self._lm_bf16 = lm_bf16
        self.list_major_copy = bool(mode)

    def _rows_by_list(self) -> Optional[torch.Tensor]:
        """The list-major copy of the bank (None unless `list_major_copy`), re-packed if the lists changed."""
        if not self.list_major_copy:
            return None
        if self._bank_by_list is None:
            if self._lm_bf16:
                self._bank_by_list = torch.empty(self.memory_features.shape, dtype=torch.bfloat16, device=self.device)
                self._lm_relerr = torch.zeros(1, dtype=torch.float32, device=self.device)
            else:
                self._bank_by_list = torch.empty_like(self.memory_features)
        if not self._by_list_valid:
            ops.ivf_pack_lists(self.memory_features, self._list_rows, self.memory_count, self._bank_by_list,
                               relerr=self._lm_relerr)
            self._by_list_valid = True
        return self._bank_by_list

    # ------------------------------------------------------------------ queries
    def _row_terms(self, location) -> Tuple[torch.Tensor, torch.Tensor]:
        """Per-row scale / bias of the combined score (hippocampal.py:282-303), cached while the fp32
        clock value, the query location and the bank contents are unchanged."""
        now32 = float(np.float32(time.time()))          # the reference subtracts in fp32 (:296)
        loc = None
        loc_key = None
        if location is not None:
            if isinstance(location, torch.Tensor) and location.is_cuda:
                loc_key = ("dev", location.data_ptr(), location._version)   # no device read just to build a cache key
            else:
                loc_key = tuple(np.asarray(location, dtype=np.float32).reshape(-1).tolist())
            loc = self._to_dev_f32(location).reshape(-1).contiguous()
        key = (now32, loc_key, self._version, self.memory_count)
        if key != self._terms_key:
            ops.row_terms(self.memory_metadata, self._inv_norm, now32, self.memory_count, self.memory_locations, loc,
                          self._scale, self._bias)
            self._terms_key = key
        return self._scale, self._bias

    def _centroid_path(self) -> bool:
        return bool(self.use_centroid_index and self._index_ready and self.memory_count > self.centroids_k)   # :259

    def _queries(self, q) -> torch.Tensor:
        q = self._to_dev_f32(q).detach()
        if q.dim() == 1:
            q = q.unsqueeze(0)
        return q.contiguous()

    def retrieve_batch(self, queries, k: int = 5, location=None, gather: bool = False, force_exact: bool = False,
                       allow_empty: bool = False):
        """Batched form of `retrieve_similar_memories`: queries [B,d] -> (rows int64 [B,k'], scores fp32 [B,k'])
        with k' = min(k, memory_count); missing results are row -1 / score -inf.  `gather=True` also returns the
        fp32 feature rows [B,k',d] (zeros for missing), replacing memory_augmented_layer.py:124-128.
        `allow_empty` (row shards, `ShardedIndex`): a query whose probed lists hold none of THIS bank's rows gets no
        result instead of the all-rows scan of :269-270, which the caller applies to the merged result."""
        self._ensure_derived()
        q = self._queries(queries)
        m = self.memory_count
        if m == 0:
            e = torch.empty(q.shape[0], 0, device=self.device)
            return (e.long(), e, e.unsqueeze(-1)) if gather else (e.long(), e)
        kk = min(int(k), m)                                                     # :306
        if not (1 <= kk <= AURA_MAX_K):
            raise ValueError(f"k={k} outside [1,{AURA_MAX_K}]")
        scale, bias = self._row_terms(location)
        if self._centroid_path() and not force_exact:
            self._ensure_lists()
            nprobe = min(self.nprobe, self.centroids_k, self._centroid_buffer_rows())   # :262
            if q.shape[0] >= ops.TC_IVF_MIN_BATCH and ops.batch_topk_supported(self.memory_features, kk):
                by_list = self._rows_by_list()
                # with a bf16 list-major shadow the bound is measured per query; eps is then the score-per-cosine unit
                # (so is it over a bf16 bank: only the query's rounding counts, the rows are exact operands)
                unit = 0.5 * self._max_strength()
                measured = self._lm_bf16 or self.memory_features.dtype == torch.bfloat16
                idx, score = ops.ivf_search_batched(self.memory_features, m, q, self.centroids, nprobe,
                                                    self._list_offsets, self._list_rows, kk, scale, bias,
                                                    eps=unit if measured else ops.TC_EPS_COS * unit,
                                                    strict=self.ivf_strict, rows_by_list=by_list,
                                                    allow_empty=allow_empty,
                                                    lm_relerr=self._lm_relerr if self._lm_bf16 else None,
                                                    measured_eps=measured)
            else:
                idx, score = ops.ivf_search(self.memory_features, m, q, self.centroids, nprobe, self._list_offsets,
                                            self._list_rows, kk, scale, bias, allow_empty=allow_empty)
        else:
            idx, score = self._exact(q, kk, scale, bias, 0.5 * self._max_strength())
        if gather:
            return idx, score, ops.gather_rows(self.memory_features, idx)
        return idx, score

    def retrieve_context(self, queries, k: int = 5, location=None):
        """What `MemoryAugmentedLayer` needs from the bank for its "concat" / "gate" injection, in one pass
        (memory_augmented_layer.py:106-130 + :185-194): top-k search of the query block, then
        context[b] = sum_j softmax(scores[b])_j * memory_features[row_j] with the reference's zero padding for
        missing results.  Returns (context fp32 [B,d], scores [B,k] zero-padded like the reference's, rows int64 [B,k]);
        the [B,k,d] feature block is never materialised."""
        idx, score = self.retrieve_batch(queries, k=k, location=location)
        if idx.shape[1] < k:                       # fewer than k memories stored: the reference pads to k (:111-112)
            pad = k - idx.shape[1]
            idx = torch.cat([idx, idx.new_full((idx.shape[0], pad), -1)], dim=1)
            score = torch.cat([score, score.new_zeros(score.shape[0], pad)], dim=1)
        ctx = ops.gather_context(self.memory_features, idx.contiguous(), score.contiguous())
        return ctx, torch.where(idx >= 0, score, torch.zeros_like(score)), idx

    def _max_strength(self) -> float:
        """Upper bound of the live strengths (the tensor-core error bound scales with it), kept on the host."""
        return self._strength_bound

    def _ensure_derived(self) -> None:
        if self._derived_stale:
            self.refresh_derived()

    def _mark_shadow_dirty(self, lo: int, hi: int) -> None:
        d = self._shadow_dirty
        if d[0] == d[1]:
            d[0], d[1] = lo, hi
        else:
            d[0], d[1] = min(d[0], lo), max(d[1], hi)

    def _shadow_rows(self) -> Optional["ops.Bf16Shadow"]:
        """The bf16 shadow of the fp32 bank, refreshed for the rows written since the last call (None unless enabled)."""
        if not self.bf16_shadow:
            return None
        if self._shadow is None:
            self._shadow = ops.Bf16Shadow(self.memory_features, n_rows=self.memory_count)
            self._shadow_dirty = [0, 0]
        lo, hi = self._shadow_dirty
        if hi > lo:
            self._shadow.refresh(self.memory_features, lo, min(hi, self.memory_features.shape[0]))
            self._shadow_dirty = [0, 0]
        return self._shadow

    def _exact(self, q: torch.Tensor, k: int, scale, bias, score_per_cos: float):
        """Exact top-k of a query block over all live rows: tcgen05 shortlist + exact fp32 re-score for blocks of
        >= TC_MIN_BATCH queries (identical results to the scan, see ops.exact_topk_batched), else the streaming scan."""
        m = self.memory_count
        if q.shape[0] >= ops.tc_min_batch(self.memory_features) and ops.batch_topk_supported(self.memory_features, k) and m >= 1024:
            shadow = self._shadow_rows() if k <= ops.TC_SHADOW_MAX_K else None
            # with the shadow the bound is measured per query; eps is then just the score-per-cosine unit
            eps = score_per_cos if shadow is not None else ops.TC_EPS_COS * score_per_cos
            return ops.exact_topk_batched(self.memory_features, q, k, scale, bias, n_rows=m, eps=eps, shadow=shadow)
        return ops.scan_topk(self.memory_features, q, k, scale, bias, n_rows=m)

    def retrieve_similar_memories(self, query_features, location=None, k: int = 5) -> List[Tuple[Union[str, int], float]]:
        """List[(memory_id, combined score)], best first (hippocampal.py:245-319)."""
        if self.memory_count == 0:
            return []                                                           # :251-252
        idx, score = self.retrieve_batch(query_features, k=k, location=location)
        # one host sync for both results: pinned staging buffers, two async copies, one stream synchronise
        kk = idx.shape[1]
        stage = self._host_stage.get(kk)
        if stage is None:
            stage = (torch.empty(kk, dtype=torch.int64).pin_memory(), torch.empty(kk, dtype=torch.float32).pin_memory())
            self._host_stage[kk] = stage
        stage[0].copy_(idx[0], non_blocking=True)
        stage[1].copy_(score[0], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        rows = stage[0].tolist()
        scores = stage[1].tolist()
        out: List[Tuple[Union[str, int], float]] = []
        for r, s in zip(rows, scores):
            if r < 0:
                break
            if not self.track_ids:
                out.append((r, s))
                continue
            owner = self._ids.owner(r)
            if owner is not None:                                               # :316
                out.append((owner, s))
        return out

    def exact_topk(self, queries, k: int = 10, gather: bool = False, defer: bool = False):
        """Pure cosine exact top-k, no temporal / strength terms (.tmp_infer_old.py:40-49).

        defer=True (query blocks on the tensor-core path only) returns (idx, score, flags, q_dev) without any host
        sync: flags[b] != 0 marks the rare query whose result is not yet certified; pass everything to
        `exact_topk_fixup` once the flags have been read.  This lets a serving loop keep several batches in flight
        (bench.py's e2e leg); defer=False does the check itself."""
        self._ensure_derived()
        q = self._queries(queries)
        m = self.memory_count
        if m == 0:                                      # empty bank: nothing to rank (retrieve_batch does the same)
            e = torch.empty(q.shape[0], 0, device=self.device)
            if defer:
                return e.long(), e, torch.zeros(q.shape[0], dtype=torch.int32, device=self.device), q
            return (e.unsqueeze(-1).expand(-1, -1, q.shape[1]).contiguous(), e, e.long()) if gather else (e.long(), e)
        kk = min(int(k), m)
        if defer:
            if not (q.shape[0] >= ops.tc_min_batch(self.memory_features) and ops.batch_topk_supported(self.memory_features, kk) and m >= 1024):
                idx, score = ops.scan_topk(self.memory_features, q, kk, self._inv_norm, None, n_rows=m)
                return idx, score, torch.zeros(q.shape[0], dtype=torch.int32, device=self.device), q
            shadow = self._shadow_rows() if kk <= ops.TC_SHADOW_MAX_K else None
            idx, score, flags = ops.exact_topk_batched(self.memory_features, q, kk, self._inv_norm, None, n_rows=m,
                                                       eps=1.0 if shadow is not None else ops.TC_EPS_COS,
                                                       defer=True, shadow=shadow)
            return idx, score, flags, q
        idx, score = self._exact(q, kk, self._inv_norm, None, 1.0)
        if gather:
            return ops.gather_rows(self.memory_features, idx), score, idx
        return idx, score

    def exact_topk_fixup(self, flags: torch.Tensor, idx: torch.Tensor, score: torch.Tensor, q_dev: torch.Tensor,
                         k: int = 10) -> int:
        """Second half of `exact_topk(..., defer=True)`: re-run the flagged queries through the exact scan, in place."""
        return ops.exact_topk_fixup(flags, idx, score, self.memory_features, q_dev, idx.shape[1], self._inv_norm, None,
                                    n_rows=self.memory_count)

    # ------------------------------------------------------------------ cognitive map
    def build_cognitive_map(self, k: int = 32) -> Tuple[torch.Tensor, torch.Tensor]:
        """All-pairs cosine similarity of the live memories, top-k neighbours per memory, self excluded
        (the O(n^2) "cognitive map" of README.md:39,64 / training_recipes.md:292-308, which the reference
        documents but never implements).  Returns (neighbour rows int64 [M,k'], similarities fp32 [M,k']),
        best first, k' = min(k, M-1); one tcgen05 GEMM with the top-k fused into its epilogue."""
        self._ensure_derived()
        m = self.memory_count
        kk = min(int(k), m - 1)
        if kk < 1:
            e = torch.empty(m, 0, device=self.device)
            self._cmap = (e.long(), e)
            return self._cmap
        nbr, sim = ops.allpairs_topk(self.memory_features, kk, self._inv_norm, n_rows=m)
        self._cmap = (nbr, sim)
        self._cmap_version = self._version
        return nbr, sim

    @property
    def cognitive_map(self) -> Dict[Tuple[Union[str, int], Union[str, int]], float]:
        """{(memory_id_i, memory_id_j): distance}, distance = 1 - cosine, over the k-NN edges - the view
        training_recipes.md:292-308 reads.  Host-side dict: meant for inspection-sized banks."""
        if getattr(self, "_cmap", None) is None or getattr(self, "_cmap_version", None) != self._version:
            self.build_cognitive_map()
        nbr, sim = (t.cpu() for t in self._cmap)
        out: Dict[Tuple[Union[str, int], Union[str, int]], float] = {}
        for i in range(nbr.shape[0]):
            a = self._ids.owner(i) if self.track_ids else i
            if a is None:
                continue
            for j, s in zip(nbr[i].tolist(), sim[i].tolist()):
                if j < 0:
                    continue
                b = self._ids.owner(j) if self.track_ids else j
                if b is not None:
                    out[(a, b)] = 1.0 - s
        return out

    def cognitive_map_edges(self, k: int = 32):
        """The k-NN edges of the cognitive map as a CSR graph for sleep-phase replay / consolidation consumers
        (TODO.md:12 "sparse graph construction (k-NN edges)"; training_recipes.md:292-308 reads the dict view):
        (indptr int64 [M+1], neighbours int64 [nnz], distance fp32 [nnz]), distance = 1 - cosine, nearest first inside
        a row.  Device tensors; `cognitive_map` is the host dict over the same edges."""
        if getattr(self, "_cmap", None) is None or getattr(self, "_cmap_version", None) != self._version \
                or self._cmap[0].shape[1] != min(int(k), max(self.memory_count - 1, 0)):
            self.build_cognitive_map(k)
        nbr, sim = self._cmap
        valid = nbr >= 0
        indptr = torch.zeros(nbr.shape[0] + 1, dtype=torch.int64, device=self.device)
        indptr[1:] = valid.sum(dim=1).cumsum(0)
        return indptr, nbr[valid], (1.0 - sim[valid])

    def replay_order(self, start_row: int, length: int, k: int = 32) -> List[int]:
        """A replay trajectory over the map for the sleep phase (hippocampal_trainer.py:327-348 replays stored items):
        greedy nearest-unvisited-neighbour walk from `start_row`, at most `length` memories (host-side, small)."""
        indptr, nbr, _ = self.cognitive_map_edges(k)
        indptr, nbr = indptr.cpu().tolist(), nbr.cpu().tolist()
        seen, out, cur = {int(start_row)}, [int(start_row)], int(start_row)
        while len(out) < length:
            nxt = next((j for j in nbr[indptr[cur]:indptr[cur + 1]] if j not in seen), None)
            if nxt is None:
                break
            seen.add(nxt); out.append(nxt); cur = nxt
        return out

    # ------------------------------------------------------------------ persistence / resume
    def refresh_derived(self) -> None:
        """Recompute everything derived from the registered buffers (inverse norms, int32 centroid ids, inverted
        lists) - call after `load_state_dict` or after writing `memory_features` / `memory_metadata` directly."""
        # every row, not only the live ones: callers of the reference set `memory_count` by hand after a resume
        self._derived_stale = False
        ops.row_inv_norms(self.memory_features, out=self._inv_norm)
        self._cid.copy_(self.memory_metadata[:, 2].to(torch.int32))
        m = self.memory_count
        self._strength_bound = float(self.memory_metadata[:m, 0].abs().max()) if m else 1.0   # one device read per resume
        self._lists_dirty = True
        self._by_list_valid = False
        self._shadow_dirty = [0, self.memory_features.shape[0]]
        self._terms_key = None
        self._version += 1

    def index_state(self) -> Dict[str, Any]:
        """The Python-side state the reference never persists (`memory_count`, `_index_ready`, knobs, id table;
        TODO.md:11, colab_l4_training.py:712-734 save only the buffers): after a resume the reference's bank is
        logically empty.  Save this dict next to `state_dict()`; `load_index_state` restores it."""
        return {
            "memory_count": int(self.memory_count), "index_ready": bool(self._index_ready),
            "centroids_k": int(self.centroids_k), "centroids_update_interval": int(self.centroids_update_interval),
            "nprobe": int(self.nprobe), "use_centroid_index": bool(self.use_centroid_index),
            "centroid_counts_len": int(self.centroid_counts.numel()),
            "id_to_idx": dict(self._ids.id_to_idx) if self.track_ids else None,
            "timestamps": {k: v.timestamp for k, v in self.episodic_memories.items()} if self.track_ids else None,
        }

    def load_index_state(self, state: Dict[str, Any]) -> None:
        """Restore `index_state()` after `load_state_dict` (buffer names are the reference's, so old checkpoints load)."""
        self.memory_count = int(state["memory_count"])
        self._index_ready = bool(state["index_ready"])
        self.centroids_k = int(state["centroids_k"])
        self.centroids_update_interval = int(state["centroids_update_interval"])
        self.nprobe = int(state["nprobe"])
        self.use_centroid_index = bool(state["use_centroid_index"])
        self._ids = IdTable()
        self.episodic_memories = {}
        if state.get("id_to_idx") is not None:
            ts = state.get("timestamps") or {}
            for mid, row in state["id_to_idx"].items():          # dict order = first-insertion order, as the reference's
                self._ids.set(mid, int(row))
                self.episodic_memories[mid] = EpisodicMemory(memory_id=mid, feature_idx=int(row),
                                                             timestamp=float(ts.get(mid, 0.0)))
        self.refresh_derived()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # the reference rebinds `centroid_counts` to a centroids_k-long tensor on every rebuild (:374): accept either length
        key = prefix + "centroid_counts"
        if key in state_dict and state_dict[key].shape != self.centroid_counts.shape:
            self.centroid_counts = torch.zeros_like(state_dict[key], device=self.device)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        # inverse norms, int32 centroid ids, inverted lists, cached score terms and the strength bound all derive from
        # the buffers just replaced: recompute lazily before the next query or write
        self._derived_stale = True
