"""GPU parity of the tcgen05 paths: aura_batch_topk (K6) and aura_allpairs_topk (K7) against the CPU oracle.

K6 bar: identical to the exact scan - rows equal except at score ties, scores within 1e-4 relative (fp32 bank);
queries the kernel cannot certify must be flagged (never silently wrong).
K7 bar (bf16 bank, fp32 accumulate): scores within 1e-2 (north_star), neighbour sets equal except near-ties.
"""
import numpy as np
import pytest
import torch

from oracle.hippo_oracle import cognitive_map_topk, exact_cosine_topk

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from aura_snn_rag_b200 import ops
    return ops


@pytest.mark.parametrize("n,d,b,k", [
    (3000, 64, 5, 10), (5000, 768, 130, 10), (20000, 768, 64, 10), (1000, 32, 3, 5), (4097, 100, 17, 7),
    (70000, 256, 300, 10), (300, 768, 200, 18),
])
def test_batch_topk_fp32_equals_oracle(n, d, b, k):
    ops = _ops()
    g = torch.Generator().manual_seed(n + d + b)
    bank = torch.randn(n, d, generator=g)
    q = bank[torch.randint(0, n, (b,), generator=g)] + 0.1 * torch.randn(b, d, generator=g)
    ref_i, ref_s = exact_cosine_topk(bank, q, k)
    rows = bank.to(DEV)
    inv = ops.row_inv_norms(rows)
    idx, sc, flags = ops.batch_topk(rows, q.to(DEV), k, inv)
    torch.cuda.synchronize()
    idx, sc, flags = idx.cpu(), sc.cpu(), flags.cpu()
    sure = flags == 0
    assert sure.float().mean() > 0.5          # certification must not be vacuous on Gaussian data
    np.testing.assert_allclose(sc[sure].numpy(), ref_s[sure].numpy(), rtol=1e-4, atol=1e-6)
    mism = (idx[sure] != ref_i[sure])
    if mism.any():   # only swaps between (near-)equal scores
        assert np.all(np.abs(sc[sure][mism].numpy() - ref_s[sure][mism].numpy()) <= 2e-6)
    # the full path (fallback for flagged queries) equals the scan path exactly
    i2, s2 = ops.exact_topk_batched(rows, q.to(DEV), k, inv)
    i3, s3 = ops.scan_topk(rows, q.to(DEV), k, inv)
    assert torch.equal(i2, i3) and torch.equal(s2, s3)


def test_batch_topk_rejects_k_beyond_the_certified_shortlist():
    ops = _ops()
    rows = torch.randn(500, 64, device=DEV)
    assert not ops.batch_topk_supported(rows, 115) and ops.batch_topk_supported(rows, 114)
    from aura_snn_rag_b200._lib import AuraLibraryError
    with pytest.raises(AuraLibraryError):
        ops.batch_topk(rows, rows[:4].contiguous(), 115, None)


@pytest.mark.parametrize("n,d,b,k,dt", [(30000, 128, 40, 19, torch.float32), (50000, 256, 150, 50, torch.float32),
                                        (40000, 768, 20, 100, torch.float32), (30000, 128, 33, 100, torch.bfloat16),
                                        (90, 64, 9, 64, torch.float32)])
def test_batch_topk_multi_round_large_k(n, d, b, k, dt):
    """k > 18: rounds of 32 candidates under a moving ceiling; the result must equal the exact scan."""
    ops = _ops()
    g = torch.Generator().manual_seed(n + k)
    bank = torch.randn(n, d, generator=g).to(dt)
    q = bank[torch.randint(0, n, (b,), generator=g)].float() + 0.3 * torch.randn(b, d, generator=g)
    rows = bank.to(DEV)
    inv = ops.row_inv_norms(rows)
    kk = min(k, n)
    i1, s1, flags = ops.batch_topk(rows, q.to(DEV), kk, inv)
    i2, s2 = ops.scan_topk(rows, q.to(DEV), kk, inv)
    sure = (flags == 0)
    assert float(sure.float().mean()) > 0.5
    assert torch.equal(i1[sure], i2[sure]) and torch.equal(s1[sure], s2[sure])
    i3, s3 = ops.exact_topk_batched(rows, q.to(DEV), kk, inv)
    assert torch.equal(i3, i2) and torch.equal(s3, s2)


def test_batch_topk_affine_terms_bf16_and_row_base():
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    n, d, b, k = 9000, 128, 40, 10
    bank = torch.randn(n, d, generator=g)
    q = torch.randn(b, d, generator=g)
    strength = 0.5 + 0.5 * torch.rand(n, generator=g)
    bias = (0.2 * torch.rand(n, generator=g) * strength).to(DEV)
    for dt in (torch.float32, torch.bfloat16):
        rows = bank.to(DEV).to(dt)
        inv = ops.row_inv_norms(rows)
        scale = 0.5 * strength.to(DEV) * inv
        i1, s1 = ops.exact_topk_batched(rows, q.to(DEV), k, scale, bias, row_base=1 << 33, eps=0.5 * ops.TC_EPS_COS)
        i2, s2 = ops.scan_topk(rows, q.to(DEV), k, scale, bias, row_base=1 << 33)
        assert torch.equal(i1, i2) and torch.equal(s1, s2)


def test_batch_topk_flags_uncertain_results():
    """Near-duplicate rows make the shortlist margin collapse: the kernel must say so."""
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    base = torch.randn(1, 64, generator=g)
    bank = base + 1e-4 * torch.randn(5000, 64, generator=g)      # 5000 rows within 1e-4 of each other
    rows = bank.to(DEV)
    inv = ops.row_inv_norms(rows)
    q = base.to(DEV).repeat(4, 1)
    _, _, flags = ops.batch_topk(rows, q, 10, inv)
    assert flags.cpu().sum() == 4
    i1, s1 = ops.exact_topk_batched(rows, q, 10, inv)
    i2, s2 = ops.scan_topk(rows, q, 10, inv)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)


@pytest.mark.parametrize("n,d,k,dt", [(2000, 64, 8, torch.bfloat16), (5000, 768, 32, torch.bfloat16),
                                      (3001, 128, 32, torch.float32), (40000, 256, 32, torch.bfloat16)])
def test_allpairs_topk_matches_oracle(n, d, k, dt):
    ops = _ops()
    g = torch.Generator().manual_seed(n + k)
    centres = torch.randn(64, d, generator=g)
    bank = (centres[torch.randint(0, 64, (n,), generator=g)] + 0.7 * torch.randn(n, d, generator=g)).to(dt)
    ref_i, ref_s = cognitive_map_topk(bank.float(), k)
    rows = bank.to(DEV)
    idx, sc = ops.allpairs_topk(rows, k)
    torch.cuda.synchronize()
    idx, sc = idx.cpu(), sc.cpu()
    tol = 1e-2 if dt == torch.bfloat16 else 3e-3      # bf16 bar of the north star; tf32 from an fp32 bank
    np.testing.assert_allclose(sc.numpy(), ref_s.numpy(), atol=tol, rtol=0)
    assert (idx != torch.arange(n).unsqueeze(1)).all()          # self excluded
    overlap = np.mean([len(set(a) & set(b)) / k for a, b in zip(idx.tolist(), ref_i.tolist())])
    assert overlap > (0.99 if dt == torch.bfloat16 else 0.97), overlap
    # sharded call: an A block in the middle of the bank gives the same rows of the result
    a0, na = n // 3, 257
    idx2, sc2 = ops.allpairs_topk(rows, k, a_first=a0, n_a=na)
    assert torch.equal(idx2.cpu(), idx[a0:a0 + na]) and torch.equal(sc2.cpu(), sc[a0:a0 + na])


@pytest.mark.parametrize("n,d,c,dt", [(20000, 128, 300, torch.float32), (9000, 768, 256, torch.float32),
                                      (30000, 256, 1000, torch.bfloat16)])
def test_tensorcore_assign_matches_exact_assign(n, d, c, dt, monkeypatch):
    """aura_kmeans_assign: tcgen05 formulation (forced) vs the exact fp32 SIMT formulation vs the oracle's cdist."""
    ops = _ops()
    g = torch.Generator().manual_seed(n + c)
    centres = torch.randn(c // 4, d, generator=g)
    bank = (centres[torch.randint(0, c // 4, (n,), generator=g)] + 0.3 * torch.randn(n, d, generator=g)).to(dt)
    cent = bank[torch.randperm(n, generator=g)[:c]].float().contiguous()
    ref = torch.argmin(torch.cdist(bank.float(), cent), dim=1)
    rows, cent_d = bank.to(DEV), cent.to(DEV)
    inv = ops.row_inv_norms(rows)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("AURA_ASSIGN_TC", mode)
        a = torch.empty(n, dtype=torch.int32, device=DEV)
        best = torch.empty(n, dtype=torch.float32, device=DEV)
        ops.kmeans_assign(rows, n, cent_d, c, a, best=best, inv_norm=inv)
        torch.cuda.synchronize()
        out[mode] = (a.cpu().long(), best.cpu())
    assert (out["0"][0] == ref).float().mean() > 0.999
    agree = (out["1"][0] == out["0"][0]).float().mean()
    assert agree > 0.999, agree
    # where they differ the two centroids are equidistant within fp32 rounding of the distance
    dist = torch.cdist(bank.float(), cent)
    bad = torch.nonzero(out["1"][0] != out["0"][0]).squeeze(-1)
    for r in bad.tolist():
        a1, a0 = out["1"][0][r], out["0"][0][r]
        assert abs(float(dist[r, a1] - dist[r, a0])) <= 1e-4 * float(dist[r, a0])


def test_tensorcore_coarse_probes_match_exact_coarse(monkeypatch):
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    c, d, b, p = 1024, 256, 200, 16
    cent = torch.randn(c, d, generator=g)
    cent[900:] = 0                                   # zeroed tail rows take part (hippocampal.py:261)
    q = cent[torch.randint(0, 900, (b,), generator=g)] + 0.5 * torch.randn(b, d, generator=g)
    dist = torch.norm(cent.unsqueeze(0) - q.unsqueeze(1), dim=2)
    ref = torch.topk(-dist, p, dim=1).indices
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("AURA_COARSE_TC", mode)
        res[mode] = ops.ivf_coarse(q.to(DEV), cent.to(DEV), p).cpu()
    # the zeroed rows tie exactly (torch.topk leaves their order open; this library takes the lower row first)
    np.testing.assert_allclose(torch.gather(dist, 1, res["0"]).numpy(), torch.gather(dist, 1, ref).numpy(), rtol=1e-6)
    assert torch.equal(res["0"][:, 1:], torch.arange(900, 900 + p - 1).expand(b, p - 1))
    assert torch.equal(res["0"][:, 0], ref[:, 0])
    # TF32 scores: same nearest centroid, same distance profile up to the rounding of 2 q.c
    assert (res["1"][:, 0] == ref[:, 0]).float().mean() > 0.99
    np.testing.assert_allclose(torch.gather(dist, 1, res["1"]).numpy(), torch.gather(dist, 1, ref).numpy(), rtol=2e-3)
    # nprobe > 32: two rounds of 32 on the tensor path
    p2 = 64
    ref2 = torch.topk(-dist, p2, dim=1).indices
    monkeypatch.setenv("AURA_COARSE_TC", "1")
    got = ops.ivf_coarse(q.to(DEV), cent.to(DEV), p2).cpu()
    assert all(len(set(r)) == p2 for r in got.tolist())                       # no repeats across rounds
    np.testing.assert_allclose(torch.gather(dist, 1, got).numpy(), torch.gather(dist, 1, ref2).numpy(), rtol=2e-3)


@pytest.mark.parametrize("n,d,c,p,b,k,dt", [(30000, 128, 64, 8, 200, 10, torch.float32),
                                            (60000, 256, 300, 16, 333, 10, torch.float32),
                                            (40000, 128, 100, 8, 128, 5, torch.bfloat16),
                                            (20000, 100, 50, 8, 100, 10, torch.float32),
                                            (50000, 128, 64, 8, 90, 100, torch.float32),
                                            (30000, 256, 40, 6, 70, 40, torch.bfloat16)])
@pytest.mark.parametrize("path", ["rows", "queries"])
def test_ivf_search_batched_equals_per_query_path(n, d, c, p, b, k, dt, path, monkeypatch):
    """List-major grouped-GEMM IVF (aura_ivf_search_batch) vs the per-query scan of the same probed lists, through both
    formulations of the fine stage: rows-as-M with one-pass selection (default for k > 18) and queries-as-M with
    register lists (default for k <= 18; k > 18 then takes rounds of 32)."""
    monkeypatch.setenv("AURA_IVF_ROWS", "1" if path == "rows" else "0")
    ops = _ops()
    g = torch.Generator().manual_seed(n + c)
    centres = torch.randn(c // 2, d, generator=g)
    bank = (centres[torch.randint(0, c // 2, (n,), generator=g)] + 0.5 * torch.randn(n, d, generator=g)).to(dt)
    rows = bank.to(DEV)
    inv = ops.row_inv_norms(rows)
    cent = torch.zeros(c + 5, d)                    # 5 zeroed tail rows, as in the reference's 256-row buffer
    cent[:c] = bank[torch.randperm(n, generator=g)[:c]].float()
    cent_d = cent.to(DEV)
    assign = torch.empty(n, dtype=torch.int32, device=DEV)
    ops.kmeans_assign(rows, n, cent_d, c, assign)
    offsets = torch.zeros(c + 6, dtype=torch.int32, device=DEV)
    lrows = torch.zeros(n, dtype=torch.int32, device=DEV)
    ops.ivf_build_lists(assign, n, c + 5, offsets, lrows)
    q = (bank[torch.randint(0, n, (b,), generator=g)].float() + 0.2 * torch.randn(b, d, generator=g)).to(DEV)
    strength = (0.5 + 0.5 * torch.rand(n, generator=g)).to(DEV)
    scale, bias = 0.5 * strength * inv, 0.1 * strength
    stats = {}
    i1, s1 = ops.ivf_search_batched(rows, n, q, cent_d, p, offsets, lrows, k, scale, bias, eps=0.5 * ops.TC_EPS_COS,
                                    stats=stats)
    i2, s2 = ops.ivf_search(rows, n, q, cent_d, p, offsets, lrows, k, scale, bias)
    torch.cuda.synchronize()
    assert stats["uncertain"] < b // 2              # the grouped pass itself must answer most queries
    assert torch.equal(s1, s2)
    assert torch.equal(i1, i2)
    # list-major resident copy (list tiles streamed by TMA instead of gathered): same answers
    packed = ops.ivf_pack_lists(rows, lrows, n, torch.empty_like(rows))
    assert torch.equal(packed, rows[lrows.long()])
    stats3 = {}
    i3, s3 = ops.ivf_search_batched(rows, n, q, cent_d, p, offsets, lrows, k, scale, bias, eps=0.5 * ops.TC_EPS_COS,
                                    stats=stats3, rows_by_list=packed)
    torch.cuda.synchronize()
    assert stats3["uncertain"] == stats["uncertain"]
    assert torch.equal(s3, s2)
    assert torch.equal(i3, i2)
    # bf16 bank with the MEASURED per-query bound (only the query is rounded; eps = unit): same answers after the fix-up,
    # and never more hand-backs than the worst-case bound gives
    if dt == torch.bfloat16:
        stats5 = {}
        i5, s5 = ops.ivf_search_batched(rows, n, q, cent_d, p, offsets, lrows, k, scale, bias, eps=0.5, stats=stats5,
                                        rows_by_list=packed, measured_eps=True)
        torch.cuda.synchronize()
        assert torch.equal(s5, s2) and torch.equal(i5, i2)
        assert stats5["handed_back"] <= stats3["handed_back"]
    # bf16 list-major SHADOW of an fp32 bank: the tensor cores read bf16 list tiles, the finish re-scores from the fp32
    # rows under a per-query measured bound - after the strict fix-up the answers are still the per-query path's
    if dt == torch.float32 and d % 8 == 0:
        relerr = torch.zeros(1, device=DEV)
        shadow = ops.ivf_pack_lists(rows, lrows, n, torch.empty(n, d, dtype=torch.bfloat16, device=DEV), relerr=relerr)
        assert torch.equal(shadow, rows[lrows.long()].to(torch.bfloat16))
        rel = ((bank.to(torch.bfloat16).double() - bank.double()).norm(dim=1) / bank.double().norm(dim=1)).max()
        assert float(rel) <= float(relerr) <= float(rel) * 1.001 + 1e-9
        stats4 = {}
        i4, s4 = ops.ivf_search_batched(rows, n, q, cent_d, p, offsets, lrows, k, scale, bias, eps=0.5, stats=stats4,
                                        rows_by_list=shadow, lm_relerr=relerr)       # eps = unit: max |scale| * ||row||
        torch.cuda.synchronize()
        assert stats4["uncertain"] < b // 2
        assert torch.equal(s4, s2)
        assert torch.equal(i4, i2)


@pytest.mark.parametrize("n,d,b,k", [(50000, 256, 200, 10), (30000, 768, 64, 34), (20000, 128, 130, 5)])
def test_bf16_shadow_shortlist_gives_the_exact_fp32_answers(n, d, b, k):
    """fp32 bank searched through its bf16 shadow (kind::f16 shortlist of 48, exact fp32 re-score from the fp32 rows):
    after the certified fix-up the results are the streaming scan's, bit for bit - also with near-duplicate rows, where
    the bf16 scores cannot separate the candidates and the flags must fire."""
    ops = _ops()
    g = torch.Generator().manual_seed(n + d)
    bank = torch.randn(n, d, generator=g)
    bank[1000:1200] = bank[999] + 1e-4 * torch.randn(200, d, generator=g)       # 200 near-duplicates of one row
    rows = bank.to(DEV)
    inv = ops.row_inv_norms(rows)
    shadow = ops.Bf16Shadow(rows)
    assert torch.equal(shadow.rows.cpu(), bank.to(torch.bfloat16))
    rel = ((bank.to(torch.bfloat16).double() - bank.double()).norm(dim=1) / bank.double().norm(dim=1)).max()
    assert float(rel) <= float(shadow.relerr) <= float(rel) * 1.001 + 1e-9          # the measured rounding-error bound
    q = bank[torch.randint(0, n, (b,), generator=g)] + 0.1 * torch.randn(b, d, generator=g)
    q[0] = bank[999]
    q = q.to(DEV)
    strength = 0.5 + 0.5 * torch.rand(n, generator=g)
    strength[999:1200] = 1.0                                                      # the near-duplicates also tie on the score terms
    strength = strength.to(DEV)
    scale, bias = 0.5 * strength * inv, 0.05 * strength
    i_ref, s_ref = ops.scan_topk(rows, q, k, scale, bias)
    _, _, flags = ops.batch_topk(rows, q, k, scale, bias, eps=0.5, shadow=shadow)     # eps = unit: max |scale| * ||row||
    assert int(flags[0]) == 1                                     # the near-duplicate query is handed back
    if k <= 10:
        assert int(flags.sum()) < b // 4                          # and with the usual margin (48 - k) most queries are certified
    stats = {}
    i_sh, s_sh = ops.exact_topk_batched(rows, q, k, scale, bias, eps=0.5, shadow=shadow, stats=stats)
    assert torch.equal(i_sh, i_ref) and torch.equal(s_sh, s_ref)
    # worst-case bound with a bare bf16 tensor: same answers, more hand-backs
    i_w, s_w = ops.exact_topk_batched(rows, q, k, scale, bias, eps=0.5 * ops.TC_EPS_COS_BF16, shadow=shadow.rows)
    assert torch.equal(i_w, i_ref) and torch.equal(s_w, s_ref)
    i_tf, s_tf = ops.exact_topk_batched(rows, q, k, scale, bias, eps=0.5 * ops.TC_EPS_COS)
    assert torch.equal(i_tf, i_ref) and torch.equal(s_tf, s_ref)


def test_bf16_shadow_follows_writes(monkeypatch):
    """HippocampalFormation(bf16_shadow=True): the shadow is refreshed for exactly the rows written since the last search."""
    import types
    import aura_snn_rag_b200.hippocampal as hmod
    monkeypatch.setattr(hmod, "time", types.SimpleNamespace(time=lambda: 1.79e9))
    g = torch.Generator().manual_seed(12)
    rows = torch.randn(6000, 128, generator=g)

    def make(shadow):
        return hmod.HippocampalFormation(n_place_cells=4, n_time_cells=2, n_grid_cells=2, max_memories=8192, feature_dim=128,
                                         use_centroid_index=False, track_ids=False, bf16_shadow=shadow)
    a, b = make(False), make(True)
    q = rows[:90] + 0.05 * torch.randn(90, 128, generator=g)
    for lo, hi in ((0, 3000), (3000, 3001), (3001, 6000)):
        for hf in (a, b):
            hf.create_episodic_memories(rows[lo:hi])
        ia, sa = a.retrieve_batch(q, 10)
        ib, sb = b.retrieve_batch(q, 10)
        assert torch.equal(ia, ib) and torch.equal(sa, sb)
        assert torch.equal(b._shadow.rows[:hi].cpu(), rows[:hi].to(torch.bfloat16))
        ea, eb = a.exact_topk(q, 7), b.exact_topk(q, 7)
        assert torch.equal(ea[0], eb[0]) and torch.equal(ea[1], eb[1])
