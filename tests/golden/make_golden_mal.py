"""Golden vectors for the batched caller of the path, from the REAL reference in the build container:

    PYTHONPATH=/root/reference:/root/reference/src python tests/golden/make_golden_mal.py

Drives the unmodified `MemoryAugmentedLayer.retrieve_memories` (memory_augmented_layer.py:86-130: a Python loop of
`retrieve_similar_memories` calls + `id_to_idx` feature look-ups) and `inject_memories` in "concat" mode (:185-190) as
UNBOUND functions on a stub that only carries what they touch (`query_proj` = identity, `hippocampus` = the unmodified
reference `HippocampalFormation` on CPU, `memory_injection`), so the rest of the layer (attention, SNN FFN) is not
constructed.  The bank stays below `centroids_k` memories, i.e. the reference answers through its exact path, whose ids
are right (the candidate-local id bug only affects the centroid path).  Output: tests/golden/mal_batch.npz.
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
from src.core.language_zone.memory_augmented_layer import MemoryAugmentedLayer  # noqa: E402

N, D, B, S, K = 200, 64, 6, 5, 7        # K > results for no query; one extra case with a 3-row bank pads with zeros
SEED = 2024


def make_inputs():
    rng = np.random.default_rng(SEED)
    rows = rng.standard_normal((N, D), dtype=np.float32)
    hidden = (rows[rng.integers(0, N, size=B)][:, None, :] + 0.3 * rng.standard_normal((B, S, D), dtype=np.float32)).astype(np.float32)
    return rows, hidden


def run(n_rows: int):
    rows, hidden = make_inputs()
    ref_mod.time = types.SimpleNamespace(time=lambda: C.T0)
    torch.manual_seed(0)
    hf = ref_mod.HippocampalFormation(2, 8, 4, 4, max_memories=256, feature_dim=D, device="cpu")
    for i in range(n_rows):
        hf.create_episodic_memory(f"m{i}", f"e{i}", torch.from_numpy(rows[i]))
    if n_rows > 50:
        hf.decay_memories(0.2)
        hf.memory_metadata[:n_rows:2, 0] *= 0.5           # two strength levels
    stub = types.SimpleNamespace(query_proj=lambda x: x, hippocampus=hf, memory_injection="concat")
    h = torch.from_numpy(hidden)
    feats, scores = MemoryAugmentedLayer.retrieve_memories(stub, h, k=K)
    out = MemoryAugmentedLayer.inject_memories(stub, h, feats, scores)
    return {"features": feats.numpy(), "scores": scores.numpy(), "injected": out.numpy(),
            "metadata": hf.memory_metadata[:n_rows].numpy().copy()}


if __name__ == "__main__":
    full, tiny = run(N), run(3)
    np.savez_compressed(os.path.join(HERE, "mal_batch.npz"), **{f"full_{k}": v for k, v in full.items()},
                        **{f"tiny_{k}": v for k, v in tiny.items()})
    print("wrote mal_batch.npz", full["scores"][0], tiny["scores"][0])
