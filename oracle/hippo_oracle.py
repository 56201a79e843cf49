"""CPU oracle for the episodic-memory retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, on CPU, the algorithm of the reference's
``HippocampalFormation`` memory bank + centroid index
(``/root/reference/src/core/hippocampal.py``) plus the two pieces the reference
only describes (pure exact cosine top-k, ``.tmp_infer_old.py:40-49``; the
all-pairs "cognitive map", ``training_recipes.md:292-308``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``aura_snn_rag_b200``) never does; it fails loudly when the CUDA library is
missing instead of falling back to anything in here.

Parity pinning
--------------
* Memory-bank / centroid-index functions: PINNED against outputs of the real
  reference imported in the build container (``tests/golden/make_golden.py``
  writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them).
  The reference's own tests hold no numeric golden vectors for this path
  (SURVEY.md section 8c), so reference outputs on seeded inputs are the contract.
* ``cognitive_map_topk``: **parity unpinned** - the reference documents the
  cognitive map but never implements it, so there is nothing to pin against.

The arithmetic deliberately goes through the same ATen CPU operators the
reference calls (``F.normalize``, ``torch.mm``, ``torch.cdist``, ``torch.topk``,
``torch.argmin``, ``torch.norm``) so that scores are bit-identical to the
reference on the same machine, not merely close.

Documented deviations from the reference ("patches"), each
switchable so the as-is behaviour can still be replayed against the goldens:

1. ``remap_candidates``: centroid-path top-k positions are candidate-local in
   the reference (``hippocampal.py:307-317``) and are looked up as if they were
   bank rows.  Patched mode maps them through ``candidates`` first.
2. ``k`` is clamped to the candidate count instead of ``memory_count``
   (``hippocampal.py:306``) in patched mode.
3. ``location`` together with the centroid path indexes
   ``memory_locations[candidates]`` (reference uses all rows and raises,
   ``hippocampal.py:287-289``).
4. ``nprobe`` and the number of rows in the ``centroids`` buffer are
   parameters (reference literals 8 and 256, ``hippocampal.py:262,114-116``).
"""

from __future__ import annotations

import time as _time
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# Score weights, hippocampal.py:301-303.
W_FEATURE = 0.5
W_SPATIAL = 0.3
W_TEMPORAL = 0.2
TEMPORAL_TAU_S = 3600.0  # hippocampal.py:297


def _as_f32(x) -> torch.Tensor:
    """numpy / tensor input handling of hippocampal.py:207-208,254-255,284-285."""
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).to(dtype=torch.float32)
    return x


class OracleHippocampus:
    """Memory bank + centroid index, CPU fp32, mirroring hippocampal.py:84-118.

    ``centroid_rows`` is the row count of the ``centroids`` buffer (256 in the
    reference); ``centroids_k`` / ``centroids_update_interval`` / ``nprobe``
    are plain attributes, mutable after construction exactly as the
    reference's tests mutate them (tests/test_hippocampal_index.py:24-25).
    """

    def __init__(
        self,
        max_memories: int = 100000,
        feature_dim: int = 768,
        spatial_dimensions: int = 2,
        use_centroid_index: bool = True,
        centroids_k: int = 256,
        centroid_rows: Optional[int] = None,
        centroids_update_interval: int = 512,
        nprobe: int = 8,
        time_fn: Callable[[], float] = _time.time,
    ) -> None:
        self.max_memories = int(max_memories)
        self.feature_dim = int(feature_dim)
        self.memory_count = 0
        self.time_fn = time_fn
        # hippocampal.py:90-99 - the three bank buffers.
        self.memory_features = torch.zeros(max_memories, feature_dim)
        self.memory_locations = torch.zeros(max_memories, spatial_dimensions)
        self.memory_metadata = torch.zeros(max_memories, 4)
        self.current_location = torch.zeros(spatial_dimensions)
        # hippocampal.py:102-103
        self.id_to_idx: Dict[str, int] = {}
        # hippocampal.py:113-118
        self.use_centroid_index = use_centroid_index
        self.centroids_k = int(centroids_k)
        self.centroids_update_interval = int(centroids_update_interval)
        self.nprobe = int(nprobe)
        rows = int(centroid_rows) if centroid_rows is not None else self.centroids_k
        self.centroids = torch.zeros(rows, feature_dim)
        self.centroid_counts = torch.zeros(rows)
        self._index_ready = False

    # ------------------------------------------------------------------ write
    def create_episodic_memory(self, memory_id: str, features, perm: Optional[torch.Tensor] = None) -> int:
        """One-shot write + online k-means step; hippocampal.py:195-243.

        Returns the bank row written.  ``perm`` is forwarded to a triggered
        ``rebuild_centroids`` (None -> ``torch.randperm`` like the reference).
        """
        # slot choice incl. the full-bank quirk (row 0 forever), :200-205
        if self.memory_count >= self.max_memories:
            row = self.memory_count % self.max_memories
        else:
            row = self.memory_count
            self.memory_count += 1
        feats = _as_f32(features)
        self.memory_features[row] = feats.detach()          # :211
        self.memory_locations[row] = self.current_location  # :212
        # fp32 storage of a ~1.8e9 timestamp is a reference quirk we keep, :215
        self.memory_metadata[row] = torch.tensor([1.0, self.time_fn(), 0.0, 0.0])
        if self.use_centroid_index and self._index_ready:
            # online nearest-centroid assign + running mean, :220-230
            live = min(self.centroids_k, self.centroids.shape[0])
            view = self.centroids[:live]
            d = torch.norm(view - feats, dim=1)
            c = torch.argmin(d)
            self.centroid_counts[c] += 1
            eta = 1.0 / self.centroid_counts[c].clamp(min=1.0)
            self.centroids[c] = (1 - eta) * self.centroids[c] + eta * feats
            self.memory_metadata[row, 2] = c
        else:
            self.memory_metadata[row, 2] = -1               # :232
        self.id_to_idx[memory_id] = row                      # :240
        # periodic full rebuild, :242-243
        if (
            self.use_centroid_index
            and self.memory_count % self.centroids_update_interval == 0
            and self.memory_count > self.centroids_k
        ):
            self.rebuild_centroids(perm=perm)
        return row

    def decay_memories(self, decay_rate: float = 0.01) -> None:
        """hippocampal.py:321-334."""
        if self.memory_count == 0:
            return
        self.memory_metadata[: self.memory_count, 0] *= (1.0 - decay_rate)

    # ---------------------------------------------------------------- rebuild
    def rebuild_centroids(self, perm: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """Seed-sample + one Lloyd step + re-assign; hippocampal.py:345-377.

        Returns the seed rows used (so a GPU build can be handed the same
        rows: CPU and CUDA ``randperm`` streams differ).
        """
        if self.memory_count == 0 or not self.use_centroid_index:
            return None
        live = self.memory_features[: self.memory_count]
        k = min(self.centroids_k, live.shape[0])              # :352
        if perm is None:
            perm = torch.randperm(live.shape[0])              # :354
        seeds = perm[:k]
        cent = live[seeds].clone()                            # :355
        assign = torch.argmin(torch.cdist(live, cent), dim=1)  # :358-359
        for c in range(k):                                    # :360-363
            members = assign == c
            if members.any():
                cent[c] = live[members].mean(dim=0)
        self.centroids[:k] = cent                             # :365
        if k < self.centroids.shape[0]:
            self.centroids[k:] = 0                            # :366-367 (tail zeroed)
        # quirk kept: the rebuild REBINDS centroid_counts to a centroids_k-long tensor (:369,:374)
        counts = torch.zeros(self.centroids_k)
        assign = torch.argmin(torch.cdist(live, self.centroids[:k]), dim=1)  # :370-371
        for c in range(k):
            counts[c] = (assign == c).sum()                   # :372-373
        self.centroid_counts = counts
        self.memory_metadata[: self.memory_count, 2] = assign.to(torch.float32)  # :376
        self._index_ready = True
        return seeds

    # ------------------------------------------------------------------ query
    def coarse_probe(self, query: torch.Tensor) -> torch.Tensor:
        """Nearest-centroid probe set; hippocampal.py:261-262.

        Scores *every* row of the centroid buffer, zeroed tail rows included
        (reference quirk), and takes ``min(nprobe, centroids_k)``.
        """
        c_d = torch.norm(self.centroids - query, dim=1)
        return torch.topk(-c_d, k=min(self.nprobe, self.centroids_k)).indices

    def candidate_rows(self, query: torch.Tensor) -> Optional[torch.Tensor]:
        """Rows whose stored centroid id is in the probe set; :258-270."""
        if not (self.use_centroid_index and self._index_ready and self.memory_count > self.centroids_k):
            return None
        probe = self.coarse_probe(query)
        cids = self.memory_metadata[: self.memory_count, 2]
        hit = torch.zeros_like(cids, dtype=torch.bool)
        for c in probe:
            hit |= cids == c
        rows = torch.nonzero(hit, as_tuple=False).squeeze(-1)
        return rows if rows.numel() > 0 else None

    def score_rows(self, query: torch.Tensor, rows: Optional[torch.Tensor], location=None,
                   patched: bool = True) -> torch.Tensor:
        """Combined score of ``rows`` (None = all live rows); hippocampal.py:272-303."""
        m = self.memory_count
        qn = F.normalize(query.unsqueeze(0), dim=1)
        feats = self.memory_features[:m] if rows is None else self.memory_features[rows]
        sim = torch.mm(qn, F.normalize(feats, dim=1).t()).squeeze(0)
        spatial = torch.zeros_like(sim)
        if location is not None:
            location = _as_f32(location)
            if rows is None or not patched:
                locs = self.memory_locations[:m]              # :287 (as-is: all rows)
            else:
                locs = self.memory_locations[rows]            # patch 3
            spatial = 1.0 / (1.0 + torch.norm(locs - location, dim=1))
        meta = self.memory_metadata[:m] if rows is None else self.memory_metadata[rows]
        ages = self.time_fn() - meta[:, 1]                    # :296 (fp32 arithmetic)
        temporal = torch.exp(-ages / TEMPORAL_TAU_S)
        return (W_FEATURE * sim + W_SPATIAL * spatial + W_TEMPORAL * temporal) * meta[:, 0]

    def retrieve_rows(self, query, location=None, k: int = 5, patched: bool = True,
                      force_exact: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """(bank rows, scores), score-descending.  ``patched=False`` replays the
        reference as-is: rows are then the reference's *candidate-local*
        positions (hippocampal.py:307, the id bug), scores are identical."""
        if self.memory_count == 0:
            return torch.empty(0, dtype=torch.long), torch.empty(0)
        query = _as_f32(query)
        rows = None if force_exact else self.candidate_rows(query)
        combined = self.score_rows(query, rows, location, patched=patched)
        kk = min(k, self.memory_count)                        # :306
        if patched:
            kk = min(kk, combined.numel())                    # patch 2
        top_s, top_i = torch.topk(combined, kk)               # :307
        if patched and rows is not None:
            top_i = rows[top_i]                               # patch 1
        return top_i, top_s

    def retrieve_similar_memories(self, query, location=None, k: int = 5, patched: bool = True,
                                  force_exact: bool = False) -> List[Tuple[str, float]]:
        """List[(memory_id, score)] like hippocampal.py:245-319."""
        rows, scores = self.retrieve_rows(query, location, k, patched, force_exact)
        inv = {v: kk for kk, v in self.id_to_idx.items()}     # :312
        out = []
        for s, r in zip(scores, rows):
            r = int(r)
            if r in inv:
                out.append((inv[r], float(s)))
        return out


# ---------------------------------------------------------------------------
# Exact-similarity path (.tmp_infer_old.py:40-49 SimpleHippocampus.retrieve)
# ---------------------------------------------------------------------------
def exact_cosine_topk(bank: torch.Tensor, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pure cosine exact top-k for a [B, d] query block against bank [M, d].

    The reference has no batched entry point (memory_augmented_layer.py:113
    loops); one mm over the block is arithmetically the same dot products.
    Returns (indices [B,k] int64, scores [B,k] fp32), score-descending.
    """
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    k = min(k, bank.shape[0])
    qn = F.normalize(queries.float(), dim=1)
    mn = F.normalize(bank.float(), dim=1)
    scores = torch.mm(qn, mn.t())
    top_s, top_i = torch.topk(scores, k, dim=1)
    return top_i, top_s


def exact_cosine_topk_f64(bank: torch.Tensor, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp64 version of the above; used to decide which fp32 orderings are ties."""
    if queries.dim() == 1:
        queries = queries.unsqueeze(0)
    k = min(k, bank.shape[0])
    qn = F.normalize(queries.double(), dim=1)
    mn = F.normalize(bank.double(), dim=1)
    scores = torch.mm(qn, mn.t())
    top_s, top_i = torch.topk(scores, k, dim=1)
    return top_i, top_s


# ---------------------------------------------------------------------------
# Cognitive map (documented, unimplemented in the reference) - PARITY UNPINNED
# ---------------------------------------------------------------------------
def cognitive_map_topk(bank: torch.Tensor, k: int = 32, block: int = 2048,
                       dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-pairs cosine similarity, top-k neighbours per row, self excluded.

    Follows the cosine of hippocampal.py:273-279 (F.normalize + mm) shaped by
    the recipe in training_recipes.md:292-308 (pairs ranked by ascending
    distance = 1 - similarity).  Returns (neighbour rows [M,k] int64,
    similarities [M,k]).  Blocked so that M x M is never materialised.
    """
    m = bank.shape[0]
    k = min(k, m - 1)
    mn = F.normalize(bank.to(dtype), dim=1)
    idx = torch.empty(m, k, dtype=torch.long)
    val = torch.empty(m, k, dtype=dtype)
    for r0 in range(0, m, block):
        r1 = min(m, r0 + block)
        s = torch.mm(mn[r0:r1], mn.t())
        s[torch.arange(r1 - r0), torch.arange(r0, r1)] = -float("inf")
        v, i = torch.topk(s, k, dim=1)
        idx[r0:r1] = i
        val[r0:r1] = v
    return idx, val


def cognitive_map_dict(ids: Sequence[str], nbr: torch.Tensor, sim: torch.Tensor) -> Dict[Tuple[str, str], float]:
    """{(id_i, id_j): distance} view the recipe reads (training_recipes.md:295-304)."""
    out: Dict[Tuple[str, str], float] = {}
    for i in range(nbr.shape[0]):
        for j, s in zip(nbr[i].tolist(), sim[i].tolist()):
            out[(ids[i], ids[j])] = 1.0 - s
    return out


# ---------------------------------------------------------------------------
# Batched caller: MemoryAugmentedLayer.retrieve_memories + inject_memories context
# ---------------------------------------------------------------------------
def retrieve_memories_batch(o: "OracleHippocampus", queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """memory_augmented_layer.py:106-130: one search per batch item (Python loop), zero-padded
    features [B,k,D] and scores [B,k].  The reference looks features up through
    ``id_to_idx``; with the patched row mapping that is ``memory_features[row]``."""
    b, d = queries.shape
    feats = torch.zeros(b, k, d)
    scores = torch.zeros(b, k)
    for i in range(b):                                       # :113
        rows, sc = o.retrieve_rows(queries[i], k=k)          # :117-120
        for j, (r, s) in enumerate(zip(rows.tolist(), sc.tolist())):
            feats[i, j] = o.memory_features[r]               # :124-128
            scores[i, j] = s
    return feats, scores


def inject_context(memory_features: torch.Tensor, memory_scores: torch.Tensor) -> torch.Tensor:
    """memory_augmented_layer.py:185-188 (and :192-193): softmax-weighted mean of the retrieved rows, [B,1,D] -> [B,D]."""
    weights = F.softmax(memory_scores, dim=-1).unsqueeze(-1)
    return (memory_features * weights).sum(dim=1)


# ---------------------------------------------------------------------------
# Host-side helpers shared by tests (CPU merge for the gloo tests, recall)
# ---------------------------------------------------------------------------
def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """k-way merge of per-shard top-k blocks: scores/ids [B, G*k'] -> [B, k].

    Order: score descending, then id ascending (the product's stated tie rule).
    """
    b = scores.shape[0]
    out_s = torch.empty(b, k, dtype=scores.dtype)
    out_i = torch.empty(b, k, dtype=ids.dtype)
    for r in range(b):
        order = sorted(range(scores.shape[1]), key=lambda j: (-float(scores[r, j]), int(ids[r, j])))[:k]
        out_s[r] = scores[r, order]
        out_i[r] = ids[r, order]
    return out_s, out_i


def recall_at_k(approx_idx: torch.Tensor, exact_idx: torch.Tensor) -> float:
    """mean |approx top-k n exact top-k| / k  (SURVEY.md 8d)."""
    hits = 0
    for a, e in zip(approx_idx.tolist(), exact_idx.tolist()):
        hits += len(set(a) & set(e))
    return hits / float(exact_idx.numel())
