"""Generate tests/golden/*.npz by running the REAL reference in the build container.

    PYTHONPATH=/root/reference:/root/reference/src python tests/golden/make_golden.py

``/root/reference`` does not exist on the GPU box, so the outputs are committed
as small fixtures; inputs are regenerated from ``cases.py`` seeds.  For every
case the script drives ``src.core.hippocampal.HippocampalFormation`` (unmodified,
device='cpu') through: N x create_episodic_memory (with the periodic rebuilds
that triggers), an optional final rebuild_centroids(), then queries through the
centroid path as-is, the exact path (index switched off), and - where the case
has locations - the exact path with a query location.

Determinism: the module's ``time`` is replaced by a settable fake clock and
``torch.randperm`` is wrapped so that every permutation prefix the reference
draws is recorded (the oracle and the CUDA build are handed the same seed rows;
CPU and CUDA randperm streams differ).
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("AURA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "src"))

import cases as C  # noqa: E402
import src.core.hippocampal as ref_mod  # noqa: E402


class Clock:
    def __init__(self):
        self.now = C.T0

    def time(self):
        return self.now


def run_case(case: C.Case) -> dict:
    rows = C.make_rows(case)
    queries = C.make_queries(case, rows)
    locs = C.make_locations(case)

    clock = Clock()
    ref_mod.time = types.SimpleNamespace(time=clock.time)  # hippocampal.py:11 `import time`

    seeds_log, seeds_at = [], []
    real_randperm = torch.randperm

    def recording_randperm(n, *a, **kw):
        p = real_randperm(n, *a, **kw)
        seeds_log.append(p[: case.centroids_k].clone().numpy())
        seeds_at.append(hf.memory_count)
        return p

    torch.manual_seed(case.seed)
    hf = ref_mod.HippocampalFormation(
        n_place_cells=8, n_time_cells=4, n_grid_cells=4,
        max_memories=case.max_memories, feature_dim=case.d, device="cpu",
    )
    hf.centroids_k = case.centroids_k
    hf.centroids_update_interval = case.interval

    torch.randperm = recording_randperm
    try:
        for i in range(case.n):
            clock.now = C.insert_time(case, i)
            if locs is not None:
                hf.update_spatial_state(torch.from_numpy(locs[i]))
            hf.create_episodic_memory(f"m{i}", f"e{i}", torch.from_numpy(rows[i]))
            if case.decay_every and (i + 1) % case.decay_every == 0:
                hf.decay_memories(0.05)
        if case.final_rebuild:
            hf.rebuild_centroids()
    finally:
        torch.randperm = real_randperm

    clock.now = C.query_time(case)
    m = hf.memory_count
    out = {
        "memory_count": np.int64(m),
        "index_ready": np.bool_(hf._index_ready),
        "centroids": hf.centroids.numpy().copy(),
        "centroid_counts": hf.centroid_counts.numpy().copy(),
        "metadata": hf.memory_metadata[:m].numpy().copy(),
        "n_rebuilds": np.int64(len(seeds_log)),
        "seeds_at": np.asarray(seeds_at, dtype=np.int64),
    }
    for j, s in enumerate(seeds_log):
        out[f"seeds_{j}"] = s.astype(np.int64)

    def pack(results, k):
        ids = np.full(k, -1, dtype=np.int64)
        sc = np.full(k, np.nan, dtype=np.float32)
        for t, (mid, s) in enumerate(results):
            ids[t] = int(mid[1:])
            sc[t] = np.float32(s)
        return ids, sc

    # centroid path as-is (ids carry the candidate-local bug, scores are right)
    asis_ids, asis_sc, exact_ids, exact_sc, loc_ids, loc_sc = [], [], [], [], [], []
    for q in queries:
        qt = torch.from_numpy(q)
        try:
            r = hf.retrieve_similar_memories(qt, k=case.k)
        except RuntimeError:  # k > candidate count (hippocampal.py:306 clamp bug)
            r = []
        a, b = pack(r, case.k)
        asis_ids.append(a); asis_sc.append(b)
        ready = hf._index_ready
        hf._index_ready = False
        a, b = pack(hf.retrieve_similar_memories(qt, k=case.k), case.k)
        exact_ids.append(a); exact_sc.append(b)
        if locs is not None:
            a, b = pack(hf.retrieve_similar_memories(qt, location=torch.from_numpy(locs[1]), k=case.k), case.k)
            loc_ids.append(a); loc_sc.append(b)
        hf._index_ready = ready
    out["asis_idnum"] = np.stack(asis_ids); out["asis_scores"] = np.stack(asis_sc)
    out["exact_idnum"] = np.stack(exact_ids); out["exact_scores"] = np.stack(exact_sc)
    if locs is not None:
        out["loc_idnum"] = np.stack(loc_ids); out["loc_scores"] = np.stack(loc_sc)
    # id -> row table at query time (stale entries after bank overflow included)
    out["id_rows"] = np.asarray([hf.id_to_idx[f"m{i}"] for i in range(case.n)], dtype=np.int64)
    return out


def main() -> None:
    for case in C.CASES:
        out = run_case(case)
        path = os.path.join(HERE, f"{case.name}.npz")
        np.savez_compressed(path, **out)
        print(f"{case.name}: rebuilds={int(out['n_rebuilds'])} m={int(out['memory_count'])} "
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
