"""Pin the oracle's batched-caller restatements (oracle.retrieve_memories_batch / inject_context) against outputs of the
real reference's MemoryAugmentedLayer.retrieve_memories + inject_memories("concat") (tests/golden/mal_batch.npz, written
by tests/golden/make_golden_mal.py)."""
import os

import numpy as np
import torch

import cases as C
from oracle.hippo_oracle import OracleHippocampus, inject_context, retrieve_memories_batch

HERE = os.path.dirname(os.path.abspath(__file__))
N, D, B, S, K, SEED = 200, 64, 6, 5, 7, 2024


def inputs():
    rng = np.random.default_rng(SEED)
    rows = rng.standard_normal((N, D), dtype=np.float32)
    hidden = (rows[rng.integers(0, N, size=B)][:, None, :] + 0.3 * rng.standard_normal((B, S, D), dtype=np.float32)).astype(np.float32)
    return rows, hidden


def build_oracle(n_rows, metadata):
    rows, hidden = inputs()
    o = OracleHippocampus(max_memories=256, feature_dim=D, time_fn=lambda: C.T0)
    o.memory_features[:n_rows] = torch.from_numpy(rows[:n_rows])
    o.memory_metadata[:n_rows] = torch.from_numpy(metadata)
    o.memory_count = n_rows
    return o, torch.from_numpy(hidden)


def test_oracle_batched_caller_matches_reference():
    gold = np.load(os.path.join(HERE, "golden", "mal_batch.npz"))
    for tag, n_rows in (("full", N), ("tiny", 3)):
        o, hidden = build_oracle(n_rows, gold[f"{tag}_metadata"])
        queries = hidden.mean(dim=1)                               # query_proj = identity in the golden run
        feats, scores = retrieve_memories_batch(o, queries, K)
        np.testing.assert_allclose(scores.numpy(), gold[f"{tag}_scores"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(feats.numpy(), gold[f"{tag}_features"], rtol=0, atol=0)
        ctx = inject_context(feats, scores)
        injected = hidden + 0.1 * ctx.unsqueeze(1)                 # memory_augmented_layer.py:189-190
        np.testing.assert_allclose(injected.numpy(), gold[f"{tag}_injected"], rtol=1e-5, atol=1e-6)
