"""Centroid-index parity at the sizes BASELINE.json's metric is quoted on (C4, and one 8-GPU shard of C5).

The CPU oracle cannot build a 10M-row index (the reference materialises M x C distances, hippocampal.py:358), so the
index is built by the CUDA path and the ORACLE'S QUERY PATH is run on it: for each sampled query the oracle
(oracle/hippo_oracle.py, `retrieve_rows` = hippocampal.py:257-307 patched) gets the GPU's centroids and a bank subset
that holds every row whose stored centroid id is in the oracle's own probe set, plus the exact top-k rows and random
distractors (rows the candidate filter must exclude).  (When a query's candidate set exceeds MAX_FULL rows - the
clustered distribution has lists of ~100k rows that every query probes - the subset holds a 40k random sample of the
candidates plus the 128 best candidates of the per-query CUDA path instead of all of them; membership of EVERY bank row in
the candidate set is still decided on the CPU from the stored centroid ids.)  Required: the same probe set, the same rows, scores within 1e-4,
hence recall@10 no lower than the oracle's centroid path at equal nprobe and seed rows.  The build itself is checked on
samples: the re-assignment against the final centroids (`argmin cdist`, :370-371) and the Lloyd mean of sampled lists
(:360-363).
"""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NOW = 1.79e9
RTOL = 1e-4
TIE_EPS = 5e-6


def _freeze_clock(monkeypatch):
    import aura_snn_rag_b200.hippocampal as hmod
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: NOW))
    return hmod


def _fill(hf, n, d, kind, n_centres, sigma, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    centres = None
    if kind == "K":
        centres = torch.nn.functional.normalize(torch.randn(n_centres, d, device=DEV, generator=g), dim=1)
    for r0 in range(0, n, 1 << 18):
        m = min(1 << 18, n - r0)
        if kind == "K":
            blk = centres[torch.randint(0, n_centres, (m,), device=DEV, generator=g)] + sigma * torch.randn(m, d, device=DEV, generator=g)
        else:
            blk = torch.randn(m, d, device=DEV, generator=g)
        hf.create_episodic_memories(blk)
    return g


MAX_FULL = 250_000


def _oracle_query_check(hf, q, idx, score, exact_idx, sample, k):
    """Run the oracle's centroid path for the sampled queries on the GPU-built index; returns (n_compared, recall_gpu,
    recall_oracle, n_probe_mismatch)."""
    from aura_snn_rag_b200 import ops
    from oracle.hippo_oracle import OracleHippocampus
    m, d = hf.memory_count, hf.memory_features.shape[1]
    cent_cpu = hf.centroids.cpu()
    cid_cpu = hf.memory_metadata[:m, 2].cpu().numpy()                    # float ids, the reference's own layout (:376)
    assert np.array_equal(cid_cpu.astype(np.int32), hf._cid[:m].cpu().numpy())
    probes_gpu = ops.ivf_coarse(q, hf.centroids, hf.nprobe).cpu()
    shell = OracleHippocampus(max_memories=1, feature_dim=d, centroids_k=hf.centroids_k,
                              centroid_rows=cent_cpu.shape[0], nprobe=hf.nprobe)
    shell.centroids = cent_cpu
    g = torch.Generator().manual_seed(5)
    compared, mismatch, rec_gpu, rec_or = 0, 0, [], []
    sc_, bi_ = hf._row_terms(None)
    qs = q[torch.tensor(sample, device=DEV)].contiguous()
    top128, _ = ops.ivf_search(hf.memory_features, m, qs, hf.centroids, hf.nprobe, hf._list_offsets, hf._list_rows, 128, sc_, bi_)
    top128 = top128.cpu().numpy()
    for si, b in enumerate(sample):
        qb = q[b].cpu()
        probe = shell.coarse_probe(qb)                                    # hippocampal.py:261-262 on the CPU
        same_probes = set(probe.tolist()) == set(probes_gpu[b].tolist())
        if not same_probes:
            # only legitimate at (near-)equal centroid distances
            dist = torch.norm(cent_cpu - qb, dim=1)
            a, c = sorted(dist[probe].tolist()), sorted(dist[probes_gpu[b]].tolist())
            np.testing.assert_allclose(a, c, rtol=1e-5)
            mismatch += 1
        cand = np.nonzero(np.isin(cid_cpu, probe.numpy().astype(np.float32)))[0]
        if cand.size == 0:
            continue
        cand_set = cand
        if cand.size > MAX_FULL:
            keep = top128[si][top128[si] >= 0]
            assert np.isin(keep, cand).all()                              # the per-query path only returns candidates
            cand_set = np.unique(np.concatenate([keep, cand[torch.randint(0, cand.size, (40_000,), generator=g).numpy()]]))
        extra = torch.randint(0, m, (6000,), generator=g).numpy()
        subset = np.unique(np.concatenate([cand_set, exact_idx[b].cpu().numpy(), extra]))
        sub_t = torch.from_numpy(subset).to(DEV)
        o = OracleHippocampus(max_memories=subset.size, feature_dim=d, centroids_k=hf.centroids_k,
                              centroid_rows=cent_cpu.shape[0], nprobe=hf.nprobe, time_fn=lambda: NOW)
        o.memory_features = hf.memory_features[sub_t].float().cpu()
        o.memory_metadata = hf.memory_metadata[sub_t].cpu()
        o.memory_locations = torch.zeros(subset.size, 2)
        o.memory_count = subset.size
        o.centroids = cent_cpu
        o._index_ready = True
        assert o.memory_count > o.centroids_k                             # the reference's gate (:259)
        prow, psc = o.retrieve_rows(qb, k=k)
        ref_rows, ref_sc = subset[prow.numpy()].tolist(), psc.tolist()
        assert set(ref_rows) <= set(cand.tolist())                        # distractors were filtered out
        ex = set(exact_idx[b, :10].tolist())
        rec_or.append(len(set(ref_rows[:10]) & ex) / 10)
        got_rows, got_sc = idx[b].tolist(), score[b].tolist()
        rec_gpu.append(len(set(got_rows[:10]) & ex) / 10)
        if not same_probes:
            continue
        n = len(ref_rows)
        assert len(got_rows) >= n and all(r == -1 for r in got_rows[n:])
        np.testing.assert_allclose(got_sc[:n], ref_sc, rtol=RTOL, atol=1e-6)
        for j in range(n):
            if got_rows[j] != ref_rows[j]:                                # a swap is only legitimate between tied scores
                assert got_rows[j] in ref_rows and abs(ref_sc[j] - ref_sc[ref_rows.index(got_rows[j])]) <= TIE_EPS * max(1.0, abs(ref_sc[j])), \
                    (b, j, got_rows, ref_rows)
        compared += 1
    return compared, float(np.mean(rec_gpu)), float(np.mean(rec_or)), mismatch


def _build_checks(hf, seeds, n_sample_rows=4096, n_sample_lists=4):
    """Samples of the build against the oracle's statements: re-assignment = argmin cdist(row, final centroids)
    (:370-371); a centroid = mean of the rows whose nearest SEED it is (:358-363)."""
    from aura_snn_rag_b200 import ops
    m, k = hf.memory_count, hf.centroids_k
    g = torch.Generator().manual_seed(9)
    rows = torch.randint(0, m, (n_sample_rows,), generator=g)
    x = hf.memory_features[rows.to(DEV)].float().cpu()
    cent = hf.centroids[:k].cpu()
    dist = torch.cdist(x, cent)
    ref = torch.argmin(dist, dim=1)
    got = hf._cid[rows.to(DEV)].cpu().long()
    agree = (ref == got).float().mean()
    assert agree >= 0.995, agree
    for r in torch.nonzero(ref != got).squeeze(-1).tolist():               # flips only between equidistant centroids
        assert abs(float(dist[r, got[r]] - dist[r, ref[r]])) <= 1e-4 * float(dist[r, ref[r]])
    # Lloyd mean of a few lists: first assignment (against the seed rows) from the CUDA assign, mean on the CPU in fp64
    seed_cent = hf.memory_features[seeds.to(DEV)].float().contiguous()
    a1 = torch.empty(m, dtype=torch.int32, device=DEV)
    ops.kmeans_assign(hf.memory_features, m, seed_cent, k, a1, inv_norm=hf._inv_norm)
    sizes = torch.bincount(a1.long(), minlength=k)
    for c in torch.randint(0, k, (n_sample_lists,), generator=g).tolist():
        members = torch.nonzero(a1 == c).squeeze(-1)
        if members.numel() == 0:
            np.testing.assert_array_equal(hf.centroids[c].cpu().numpy(), seed_cent[c].cpu().numpy())   # keeps its seed (:360-363)
            continue
        mean = hf.memory_features[members].double().mean(dim=0).float().cpu()
        np.testing.assert_allclose(hf.centroids[c].cpu().numpy(), mean.numpy(), rtol=1e-4, atol=1e-5)
    assert int(sizes.sum()) == m


@pytest.mark.parametrize("kind", ["K", "G"])
def test_c4_scale_ivf_matches_oracle(kind, monkeypatch):
    """BASELINE config 4: 10M x 1024 fp32, 4096 centroids, nprobe 32, batch 4096, k = 10."""
    hmod = _freeze_clock(monkeypatch)
    n, d, c, p, b, k = 10_000_000, 1024, 4096, 32, 4096, 10
    hf = hmod.HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=n, feature_dim=d,
                                   device=DEV, centroids_k=c, nprobe=p, track_ids=False)
    hf.centroids_update_interval = 1 << 40
    g = _fill(hf, n, d, kind, 1024, 0.05, 1234)
    seeds = torch.randperm(n, device=DEV, generator=g)[:c]
    hf.rebuild_centroids(seed_rows=seeds)
    _build_checks(hf, seeds)
    pick = torch.randint(0, n, (b,), device=DEV, generator=g)
    noise = 0.005 if kind == "K" else 0.1
    q = hf.memory_features[pick] + noise * torch.randn(b, d, device=DEV, generator=g)
    idx, score = hf.retrieve_batch(q, k)                    # TF32 coarse shortlist + exact finish, list-major fine stage
    ex_idx, ex_score = hf.retrieve_batch(q, k, force_exact=True)
    torch.cuda.synchronize()
    assert bool((score[:, :-1] >= score[:, 1:]).all())
    sample = torch.randperm(b, generator=torch.Generator().manual_seed(3))[:72].tolist()
    compared, rec_gpu, rec_or, mismatch = _oracle_query_check(hf, q, idx, score, ex_idx, sample, k)
    assert compared >= 48 and mismatch <= 3, (compared, mismatch)
    assert rec_gpu >= rec_or - 1e-9, (rec_gpu, rec_or)
    if kind == "K":
        assert rec_gpu >= 0.95, rec_gpu
    # the per-query path (one launch per query, fp32 scan of the probed lists) gives the same answers
    from aura_snn_rag_b200 import ops
    sc_, bi_ = hf._row_terms(None)
    i2, s2 = ops.ivf_search(hf.memory_features, n, q[:48].contiguous(), hf.centroids, p, hf._list_offsets, hf._list_rows, k, sc_, bi_)
    assert torch.equal(i2, idx[:48]) and torch.equal(s2, score[:48])
    # list-major resident copy: same answers
    hf.list_major_copy = True
    i3, s3 = hf.retrieve_batch(q, k)
    assert torch.equal(i3, idx) and torch.equal(s3, score)
    # bf16 list-major shadow (half the list bytes; exact fp32 re-score + per-query measured certificate): same answers
    hf.set_list_major_copy("bf16")
    torch.cuda.empty_cache()
    i4, s4 = hf.retrieve_batch(q, k)
    assert hf._bank_by_list.dtype == torch.bfloat16
    assert torch.equal(i4, idx) and torch.equal(s4, score)


def test_c5_shard_ivf_matches_oracle(monkeypatch):
    """One shard of BASELINE config 5 at 8 GPUs: 12.5M x 768 bf16, 16384 centroids, nprobe 64, k = 100, batch 4096."""
    hmod = _freeze_clock(monkeypatch)
    n, d, c, p, b, k = 12_500_000, 768, 16384, 64, 4096, 100
    hf = hmod.HippocampalFormation(n_place_cells=8, n_time_cells=4, n_grid_cells=4, max_memories=n, feature_dim=d,
                                   device=DEV, centroids_k=c, nprobe=p, bank_dtype=torch.bfloat16, track_ids=False)
    hf.centroids_update_interval = 1 << 40
    g = _fill(hf, n, d, "K", 8192, 0.05, 1234)
    seeds = torch.randperm(n, device=DEV, generator=g)[:c]
    hf.rebuild_centroids(seed_rows=seeds)
    _build_checks(hf, seeds, n_sample_rows=2048)
    pick = torch.randint(0, n, (b,), device=DEV, generator=g)
    q = hf.memory_features[pick].float() + 0.005 * torch.randn(b, d, device=DEV, generator=g)
    idx, score = hf.retrieve_batch(q, k)                    # strict: uncertified queries re-run through the per-query scan
    nq = 256
    ex_idx, _ = hf.retrieve_batch(q[:nq].contiguous(), k, force_exact=True)
    torch.cuda.synchronize()
    assert bool((idx[:, 0] == pick).all())
    sample = torch.randperm(nq, generator=torch.Generator().manual_seed(3))[:64].tolist()
    compared, rec_gpu, rec_or, mismatch = _oracle_query_check(hf, q, idx, score, ex_idx, sample, k)
    assert compared >= 48 and mismatch <= 3, (compared, mismatch)
    assert rec_gpu >= rec_or - 1e-9 and rec_gpu >= 0.9, (rec_gpu, rec_or)
    # relaxed mode keeps the exactly re-scored shortlist for uncertified queries: same scores wherever rows agree
    hf.ivf_strict = False
    i2, s2 = hf.retrieve_batch(q, k)
    overlap = (i2.unsqueeze(2) == idx.unsqueeze(1)).any(dim=2).float().mean()
    assert float(overlap) >= 0.999, float(overlap)
