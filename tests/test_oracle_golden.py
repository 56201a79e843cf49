"""Pin the CPU oracle against outputs of the real reference (tests/golden/*.npz).

The goldens were produced by tests/golden/make_golden.py from the unmodified
reference (src/core/hippocampal.py).  The oracle must reproduce the final index
state and every query result: as-is mode replays the reference including its
candidate-local id bug; patched mode must give the same SCORES with remapped rows.
"""
import numpy as np
import pytest
import torch

import cases as C
from golden_replay import load, replay_writes
from oracle.hippo_oracle import OracleHippocampus

RTOL, ATOL = 1e-5, 1e-6   # other host CPUs may take different sgemm code paths


class _Clock:
    now = C.T0

    def __call__(self):
        return self.now


def build_oracle(case, gold):
    clock = _Clock()
    o = OracleHippocampus(max_memories=case.max_memories, feature_dim=case.d, centroids_k=case.centroids_k,
                          centroid_rows=256, centroids_update_interval=case.interval, time_fn=clock)

    def set_time(t):
        clock.now = t

    def create(i, row, next_seeds):
        will_rebuild = (o.memory_count + (0 if o.memory_count >= o.max_memories else 1))
        trig = (will_rebuild % o.centroids_update_interval == 0) and will_rebuild > o.centroids_k
        o.create_episodic_memory(f"m{i}", row, perm=next_seeds(will_rebuild) if trig else None)

    def set_loc(l):
        o.current_location = l

    rows, queries, locs = replay_writes(case, gold, o, set_time, create,
                                        lambda s: o.rebuild_centroids(perm=s), o.decay_memories, set_loc)
    return o, rows, queries, locs


@pytest.mark.parametrize("case", C.CASES, ids=lambda c: c.name)
def test_oracle_matches_reference_state(case):
    gold = load(case.name)
    o, *_ = build_oracle(case, gold)
    m = int(gold["memory_count"])
    assert o.memory_count == m
    assert o._index_ready == bool(gold["index_ready"])
    np.testing.assert_allclose(o.centroids.numpy(), gold["centroids"], rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(o.centroid_counts.numpy(), gold["centroid_counts"])
    meta = o.memory_metadata[:m].numpy()
    np.testing.assert_allclose(meta[:, 0], gold["metadata"][:, 0], rtol=1e-6)      # strength
    np.testing.assert_array_equal(meta[:, 1], gold["metadata"][:, 1])               # fp32 timestamps
    assert (meta[:, 2] == gold["metadata"][:, 2]).mean() >= 0.999                   # centroid ids
    np.testing.assert_array_equal(np.asarray([o.id_to_idx[f"m{i}"] for i in range(case.n)]), gold["id_rows"])


@pytest.mark.parametrize("case", C.CASES, ids=lambda c: c.name)
def test_oracle_matches_reference_queries(case):
    gold = load(case.name)
    o, rows, queries, locs = build_oracle(case, gold)
    id_rows = gold["id_rows"]
    for qi, q in enumerate(queries):
        qt = torch.from_numpy(q)
        # centroid path as-is: same (buggy) ids, same scores
        g_ids, g_sc = gold["asis_idnum"][qi], gold["asis_scores"][qi]
        n_ret = int((g_ids >= 0).sum())
        if n_ret:
            res = o.retrieve_similar_memories(qt, k=case.k, patched=False)
            assert len(res) == n_ret
            np.testing.assert_allclose([s for _, s in res], g_sc[:n_ret], rtol=RTOL, atol=ATOL)
            assert [int(i[1:]) for i, _ in res] == g_ids[:n_ret].tolist()
            # patched: identical score list, rows remapped through the candidate set
            prow, psc = o.retrieve_rows(qt, k=case.k, patched=True)
            np.testing.assert_allclose(psc.numpy()[:n_ret], g_sc[:n_ret], rtol=RTOL, atol=ATOL)
            cand = o.candidate_rows(qt)
            if cand is not None:
                local = torch.as_tensor([id_rows[i] for i in g_ids[:n_ret]])
                assert torch.equal(cand[local], prow[:n_ret])
        # exact path
        g_ids, g_sc = gold["exact_idnum"][qi], gold["exact_scores"][qi]
        res = o.retrieve_similar_memories(qt, k=case.k, force_exact=True)
        n_ret = int((g_ids >= 0).sum())
        assert len(res) == n_ret
        np.testing.assert_allclose([s for _, s in res], g_sc[:n_ret], rtol=RTOL, atol=ATOL)
        assert [int(i[1:]) for i, _ in res] == g_ids[:n_ret].tolist()
        if locs is not None:
            g_ids, g_sc = gold["loc_idnum"][qi], gold["loc_scores"][qi]
            res = o.retrieve_similar_memories(qt, location=torch.from_numpy(locs[1]), k=case.k, force_exact=True)
            np.testing.assert_allclose([s for _, s in res], g_sc[: len(res)], rtol=RTOL, atol=ATOL)
            assert [int(i[1:]) for i, _ in res] == g_ids[: len(res)].tolist()


def test_reference_structural_assertions_hold_for_oracle():
    """The reference's own (structural) assertions, tests/test_hippocampal_index.py:41-51,71-74
    and tests/test_hippocampal_formation.py:75-79, replayed on the oracle."""
    case = C.CASE_BY_NAME["idx_d4_n20"]
    o, *_ = build_oracle(case, load(case.name))
    assert o._index_ready and o.memory_count == 20
    res = o.retrieve_similar_memories(torch.tensor([1.0, 0.0, 0.0, 0.0]), k=5)
    assert len(res) == 5 and all(int(i[1:]) < 10 for i, _ in res)   # cluster A = first 10 rows
    small = OracleHippocampus(max_memories=50, feature_dim=4)
    for i in range(3):
        small.create_episodic_memory(f"S{i}", torch.tensor([float(i == 0), float(i == 1), 0.0, 0.0]))
    assert small.memory_count == 3 and not small._index_ready
    assert len(small.retrieve_similar_memories(torch.tensor([1.0, 0.0, 0.0, 0.0]), k=2)) == 2
    case = C.CASE_BY_NAME["bank_d64_n5"]
    o, rows, *_ = build_oracle(case, load(case.name))
    assert o.retrieve_similar_memories(torch.from_numpy(rows[0]), k=1)[0][0] == "m0"
