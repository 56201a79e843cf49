"""Round-2 parity cases: the tensor-core coarse stage together with the batched fine stage against the oracle, the
cooperative write-run kernel at the BASELINE config 4 centroid shape, device placement, and lazy refresh after
`load_state_dict`."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
NOW = 1.79e9
TIE_EPS = 5e-6


def _same_topk(ids, sc, ref_ids, ref_sc, rtol=1e-4):
    n = len(ref_ids)
    assert len(ids) >= n and all(i == -1 for i in ids[n:]), (ids, ref_ids)
    np.testing.assert_allclose(sc[:n], ref_sc, rtol=rtol, atol=1e-6)
    for j in range(n):
        if ids[j] != ref_ids[j]:
            assert ids[j] in ref_ids and abs(ref_sc[j] - ref_sc[ref_ids.index(ids[j])]) <= TIE_EPS * max(1.0, abs(ref_sc[j])), \
                (j, ids, ref_ids)


@pytest.mark.parametrize("n,d,c,p,b,k,dt", [(60000, 128, 1024, 16, 256, 10, torch.float32),
                                            (80000, 96, 2048, 40, 160, 10, torch.float32),      # nprobe > 32: 2 shortlist rounds
                                            (50000, 128, 1024, 8, 300, 5, torch.bfloat16),
                                            (60000, 128, 1024, 16, 256, 60, torch.float32)])     # k > 18: rows-as-M one-pass path
def test_batched_ivf_with_tensorcore_coarse_matches_oracle(n, d, c, p, b, k, dt, monkeypatch):
    """B * C >= 2.5e5: the coarse stage is the TF32 GEMM shortlist + exact fp32 finish and the fine stage the list-major
    tensor-core pass.  Every query is compared with the ORACLE's centroid path (hippocampal.py:257-307, patched) run on
    the index state the CUDA build produced: same probes, same rows, scores within 1e-4."""
    import aura_snn_rag_b200.hippocampal as hmod
    from aura_snn_rag_b200 import ops
    from oracle.hippo_oracle import OracleHippocampus
    assert b >= 64 and b * (c + 8) >= 2.5e5
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: NOW))
    g = torch.Generator().manual_seed(n + c)
    centres = torch.randn(c // 4, d, generator=g)
    rows = centres[torch.randint(0, c // 4, (n,), generator=g)] + 0.6 * torch.randn(n, d, generator=g)
    hf = hmod.HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=n, feature_dim=d,
                                   centroids_k=c, centroid_rows=c + 8, nprobe=p, bank_dtype=dt, track_ids=False)
    hf.centroids_update_interval = 1 << 40
    hf.create_episodic_memories(rows)
    hf.rebuild_centroids(seed_rows=torch.randperm(n, generator=g)[:c])
    hf.decay_memories(0.1)
    hf.memory_metadata[: n // 2, 0] *= 0.7                                # two strength levels: ranking != cosine ranking
    hf._version += 1
    q = rows[torch.randint(0, n, (b,), generator=g)] + 0.3 * torch.randn(b, d, generator=g)
    idx, sc = hf.retrieve_batch(q, k)
    probes = ops.ivf_coarse(q.cuda(), hf.centroids, p).cpu()
    o = OracleHippocampus(max_memories=n, feature_dim=d, centroids_k=c, centroid_rows=c + 8, nprobe=p, time_fn=lambda: NOW)
    o.memory_features = hf.memory_features.float().cpu()
    o.memory_metadata = hf.memory_metadata.cpu()
    o.memory_count = n
    o.centroids = hf.centroids.cpu()
    o._index_ready = True
    idx, sc = idx.cpu(), sc.cpu()
    same_probes = 0
    for j in range(b):
        ref_p = o.coarse_probe(q[j])
        if set(ref_p.tolist()) != set(probes[j].tolist()):
            dist = torch.norm(o.centroids - q[j], dim=1)                  # only at (near-)equal centroid distances
            np.testing.assert_allclose(sorted(dist[ref_p].tolist()), sorted(dist[probes[j]].tolist()), rtol=1e-5)
            continue
        same_probes += 1
        prow, psc = o.retrieve_rows(q[j], k=k)
        _same_topk(idx[j].tolist(), sc[j].tolist(), prow.tolist(), psc.tolist(), rtol=1e-4)
    assert same_probes >= 0.97 * b, same_probes
    # and the per-query path (exact fp32 coarse in difference form, streaming scan) agrees bit for bit
    sc_, bi_ = hf._row_terms(None)
    i2, s2, p2 = ops.ivf_search(hf.memory_features, n, q[:40].cuda(), hf.centroids, p, hf._list_offsets, hf._list_rows, k,
                                sc_, bi_, return_probes=True)
    assert torch.equal(p2.cpu(), probes[:40])
    assert torch.equal(i2.cpu(), idx[:40]) and torch.equal(s2.cpu(), sc[:40])


def test_write_run_kernel_at_c4_centroid_shape(monkeypatch):
    """`online_assign_run_kernel` (one cooperative launch for a run of writes, centroid slices stationary in shared
    memory) at the BASELINE config 4 shape - 4096 centroids x 1024, 113 KB of shared memory per SM - against the
    one-launch-per-write kernel: bit-identical centroids, counts and assignments (hippocampal.py:218-230)."""
    import aura_snn_rag_b200.hippocampal as hmod
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: NOW))
    d, c, n0, n_new = 1024, 4096, 9000, 1200
    g = torch.Generator().manual_seed(2)
    centres = torch.nn.functional.normalize(torch.randn(512, d, generator=g), dim=1)
    rows = centres[torch.randint(0, 512, (n0 + n_new,), generator=g)] + 0.05 * torch.randn(n0 + n_new, d, generator=g)
    seeds = torch.randperm(n0, generator=g)[:c]

    def make():
        hf = hmod.HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=n0 + n_new,
                                       feature_dim=d, centroids_k=c, nprobe=8, track_ids=False)
        hf.centroids_update_interval = 1 << 40
        hf.create_episodic_memories(rows[:n0])
        hf.rebuild_centroids(seed_rows=seeds)
        return hf

    a, b = make(), make()
    assert torch.equal(a.centroids, b.centroids)
    a.create_episodic_memories(rows[n0:])                                  # one cooperative launch
    for i in range(n0, n0 + n_new):
        b.create_episodic_memory(str(i), "e", rows[i])                     # one launch per write
    m = n0 + n_new
    assert a.memory_count == b.memory_count == m
    assert torch.equal(a._cid[:m], b._cid[:m])
    assert torch.equal(a.memory_metadata[:m], b.memory_metadata[:m])
    assert torch.equal(a.centroid_counts, b.centroid_counts)
    assert torch.equal(a.centroids, b.centroids)
    # and the kernel against the oracle's statements on the CPU for a short run (before fp32 order effects can pile up)
    from oracle.hippo_oracle import OracleHippocampus
    ref = make()
    o = OracleHippocampus(max_memories=m, feature_dim=d, centroids_k=c, centroid_rows=c, time_fn=lambda: NOW)
    o.memory_features[:n0] = rows[:n0]
    o.memory_count = n0
    o.centroids = ref.centroids.cpu().clone()
    o.centroid_counts = ref.centroid_counts.cpu().clone()
    o._index_ready = True
    o.centroids_update_interval = 1 << 40
    for i in range(n0, n0 + 64):
        o.create_episodic_memory(str(i), rows[i])
    ref.create_episodic_memories(rows[n0:n0 + 64])
    got = ref._cid[n0:n0 + 64].cpu().long()
    want = o.memory_metadata[n0:n0 + 64, 2].long()
    assert (got == want).float().mean() >= 0.95                            # near-equidistant centroids may flip
    if bool((got == want).all()):
        np.testing.assert_allclose(ref.centroids.cpu().numpy(), o.centroids.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(ref.centroid_counts.cpu().numpy(), o.centroid_counts.numpy())


def test_module_on_a_non_current_device():
    """The reference class works on any device index (hippocampal.py:50-53); every ops entry makes the tensors' device
    current for its launches, streams and scratch buffers."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from aura_snn_rag_b200 import HippocampalFormation
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(1)
    rows = torch.randn(5000, 64, generator=g)
    out = {}
    for dev in ("cuda:0", "cuda:1"):
        hf = HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=8192, feature_dim=64,
                                  device=dev, centroids_k=32, nprobe=4, track_ids=False)
        hf.centroids_update_interval = 1 << 40
        hf.create_episodic_memories(rows)
        hf.rebuild_centroids(seed_rows=torch.arange(0, 5000, 150)[:32])
        out[dev] = (hf.exact_topk(rows[:70] + 0.01, 10), hf.retrieve_batch(rows[:70] + 0.01, 10), hf.retrieve_batch(rows[:3], 5))
        assert out[dev][0][0].device == torch.device(dev)
    assert torch.cuda.current_device() == 0
    for a, b in zip(out["cuda:0"], out["cuda:1"]):
        assert torch.equal(a[0].cpu(), b[0].cpu()) and torch.equal(a[1].cpu(), b[1].cpu())


def test_load_state_dict_refreshes_derived_state_lazily(monkeypatch):
    """`load_state_dict` alone (no `load_index_state`) must not leave stale inverse norms / lists behind: a caller that
    restores the reference's buffers and sets `memory_count` by hand gets correct scores."""
    import aura_snn_rag_b200.hippocampal as hmod
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: NOW))
    g = torch.Generator().manual_seed(8)
    rows = torch.randn(3000, 48, generator=g)

    def make():
        hf = hmod.HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=4096, feature_dim=48,
                                       centroids_k=16, nprobe=4, track_ids=False)
        hf.centroids_update_interval = 1 << 40
        return hf

    a = make()
    a.create_episodic_memories(rows)
    a.rebuild_centroids(seed_rows=torch.arange(0, 3000, 180)[:16])
    a.decay_memories(0.3)
    b = make()
    b.load_state_dict(a.state_dict(), strict=True)
    b.memory_count, b._index_ready = a.memory_count, True                 # what a reference user restores by hand
    q = rows[:80] + 0.05
    for kw in ({}, {"force_exact": True}):
        ia, sa = a.retrieve_batch(q, 7, **kw)
        ib, sb = b.retrieve_batch(q, 7, **kw)
        assert torch.equal(ia, ib) and torch.equal(sa, sb)
    ia, sa = a.retrieve_batch(q[:2], 7)
    ib, sb = b.retrieve_batch(q[:2], 7)
    assert torch.equal(ia, ib) and torch.equal(sa, sb)
    assert abs(b._max_strength() - 0.7) < 1e-6
    # exact_topk on an empty bank returns empty results like retrieve_batch
    e = make()
    i0, s0 = e.exact_topk(q[:3], 5)
    assert i0.shape == (3, 0) and s0.shape == (3, 0)


def test_batched_caller_and_injection_context_match_the_reference(monkeypatch):
    """`retrieve_batch(gather=True)` and the fused `retrieve_context` against outputs of the
    REAL reference's MemoryAugmentedLayer.retrieve_memories + inject_memories("concat") (tests/golden/mal_batch.npz) and
    against the oracle's restatement of that loop - not against this class's own per-item path."""
    import os
    import aura_snn_rag_b200.hippocampal as hmod
    from aura_snn_rag_b200 import memory_api as api
    from test_oracle_mal_golden import B, D, K, N, build_oracle, inputs
    import cases as C
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: C.T0))
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mal_batch.npz"))
    rows, hidden = inputs()
    for tag, n_rows in (("full", N), ("tiny", 3)):
        hf = hmod.HippocampalFormation(2, 8, 4, 4, max_memories=256, feature_dim=D, device="cuda")
        hf.create_episodic_memories(torch.from_numpy(rows[:n_rows]), [f"m{i}" for i in range(n_rows)])
        hf.memory_metadata[:n_rows] = torch.from_numpy(gold[f"{tag}_metadata"]).cuda()      # strengths after decay
        hf._version += 1
        q = torch.from_numpy(hidden).mean(dim=1).cuda()
        feats, scores = api.retrieve_memories(hf, q, k=K)                                   # [B,K,D], [B,K], zero-padded
        np.testing.assert_allclose(scores.cpu().numpy(), gold[f"{tag}_scores"], rtol=1e-4, atol=1e-6)
        np.testing.assert_array_equal(feats.cpu().numpy(), gold[f"{tag}_features"])
        ctx, sc2, idx = hf.retrieve_context(q, k=K)
        injected = torch.from_numpy(hidden).cuda() + 0.1 * ctx.unsqueeze(1)                   # memory_augmented_layer.py:189-190
        np.testing.assert_allclose(injected.cpu().numpy(), gold[f"{tag}_injected"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(sc2.cpu().numpy(), gold[f"{tag}_scores"], rtol=1e-4, atol=1e-6)
        # and the oracle's restatement of the same loop on the same state
        from oracle.hippo_oracle import inject_context, retrieve_memories_batch
        o, h = build_oracle(n_rows, gold[f"{tag}_metadata"])
        f_o, s_o = retrieve_memories_batch(o, h.mean(dim=1), K)
        np.testing.assert_allclose(ctx.cpu().numpy(), inject_context(f_o, s_o).numpy(), rtol=1e-4, atol=1e-6)
    # a larger block through the tensor-core path (B >= 3) with the centroid index on, against the oracle loop
    g = torch.Generator().manual_seed(5)
    n, d, k = 5000, 128, 6
    bank = torch.randn(n, d, generator=g)
    hf = hmod.HippocampalFormation(2, 8, 4, 4, max_memories=8192, feature_dim=d, device="cuda", centroids_k=32, nprobe=4, track_ids=False)
    hf.centroids_update_interval = 1 << 40
    hf.create_episodic_memories(bank)
    hf.rebuild_centroids(seed_rows=torch.arange(0, n, 150)[:32])
    from oracle.hippo_oracle import OracleHippocampus, inject_context, retrieve_memories_batch
    o = OracleHippocampus(max_memories=8192, feature_dim=d, centroids_k=32, centroid_rows=256, nprobe=4, time_fn=lambda: C.T0)
    o.memory_features[:n] = bank
    o.memory_metadata[:n] = hf.memory_metadata[:n].cpu()
    o.memory_count = n
    o.centroids = hf.centroids.cpu()
    o._index_ready = True
    q = bank[:40] + 0.2 * torch.randn(40, d, generator=g)
    ctx, sc, idx = hf.retrieve_context(q, k=k)
    f_o, s_o = retrieve_memories_batch(o, q, k)
    np.testing.assert_allclose(sc.cpu().numpy(), s_o.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(ctx.cpu().numpy(), inject_context(f_o, s_o).numpy(), rtol=1e-4, atol=1e-5)


def test_cognitive_map_edges_and_exact_fp32_scores():
    """fp32 bank: the map's neighbour scores are exact fp32 cosines (re-scored, 1e-4 bar of the north star; the TF32
    products only pick the candidates).  CSR edge view + replay walk for the sleep-phase consumer."""
    from aura_snn_rag_b200 import HippocampalFormation
    from oracle.hippo_oracle import cognitive_map_topk
    g = torch.Generator().manual_seed(33)
    n, d, k = 3000, 96, 12
    centres = torch.randn(30, d, generator=g)
    rows = centres[torch.randint(0, 30, (n,), generator=g)] + 0.5 * torch.randn(n, d, generator=g)
    hf = HippocampalFormation(max_memories=4096, feature_dim=d, n_place_cells=4, n_time_cells=2, n_grid_cells=2,
                              use_centroid_index=False, track_ids=False)
    hf.create_episodic_memories(rows)
    nbr, sim = hf.build_cognitive_map(k)
    ref_i, ref_s = cognitive_map_topk(rows, k)
    np.testing.assert_allclose(sim.cpu().numpy(), ref_s.numpy(), rtol=1e-4, atol=2e-6)
    assert np.mean([len(set(a) & set(b)) / k for a, b in zip(nbr.cpu().tolist(), ref_i.tolist())]) > 0.995
    assert bool((sim[:, :-1] >= sim[:, 1:]).all())
    indptr, nb, dist = hf.cognitive_map_edges(k)
    assert indptr.tolist() == list(range(0, n * k + 1, k)) and nb.numel() == n * k
    np.testing.assert_allclose(dist.cpu().numpy(), (1.0 - sim).reshape(-1).cpu().numpy())
    walk = hf.replay_order(5, 20, k)
    assert walk[0] == 5 and len(walk) == len(set(walk)) == 20
    for a, b in zip(walk[:-1], walk[1:]):
        assert b in nbr[a].tolist()


@pytest.mark.parametrize("n,d,b,k,dt,shadow", [(400_000, 64, 300, 10, torch.float32, True),      # 98 lists per row: best sample of each
                                               (400_000, 64, 300, 10, torch.float32, False),
                                               (120_000, 64, 1300, 10, torch.bfloat16, False),   # 26 lists per row: second best of each
                                               (300_000, 128, 520, 18, torch.float32, True)])
def test_sampled_start_threshold_keeps_exact_results(n, d, b, k, dt, shadow):
    """Banks large enough that every list of the dense kernel starts from the sampled threshold (a few tiles of its own
    column range scored in a cheap mode first, gemm_topk.cu): the threshold must be a true lower bound of every query's
    L-th best score, i.e. results after the certified fix-up are still the exact scan's, bit for bit - including queries
    whose best rows all sit in ONE list's range (clustered by row position) and rows with extreme scale / bias terms."""
    from aura_snn_rag_b200 import ops
    g = torch.Generator().manual_seed(n + b)
    bank = torch.randn(n, d, generator=g)
    # a block of rows near query 0 at the very end of the bank: all of its neighbours fall into the last list's range
    bank[-300:] = bank[-301] + 0.05 * torch.randn(300, d, generator=g)
    rows = bank.to(dt).cuda()
    inv = ops.row_inv_norms(rows)
    q = bank[torch.randint(0, n, (b,), generator=g)] + 0.2 * torch.randn(b, d, generator=g)
    q[0] = bank[-301]
    q = q.cuda()
    strength = 0.5 + 0.5 * torch.rand(n, generator=g)
    strength[::1000] = 5.0                                              # a few rows whose terms dominate their tile's maximum
    strength = strength.cuda()
    scale, bias = 0.5 * strength * inv, 0.05 * strength
    sh = ops.Bf16Shadow(rows) if shadow else None
    eps = 2.5 if shadow else 2.5 * (ops.TC_EPS_COS_BF16 if dt == torch.bfloat16 else ops.TC_EPS_COS)   # unit: max |scale| * ||row||
    stats = {}
    i1, s1 = ops.exact_topk_batched(rows, q, k, scale, bias, eps=eps, stats=stats, shadow=sh)
    i2, s2 = ops.scan_topk(rows, q, k, scale, bias)
    torch.cuda.synchronize()
    assert stats["uncertain"] < b // 4
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
