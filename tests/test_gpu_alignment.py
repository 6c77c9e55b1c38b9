"""GPU parity of the alignment scorer (tcgen05 similarity GEMM + streaming top-k / CSLS / rank) against the oracle.

Two operand precisions (north_star: "bf16/tf32 GEMM"):
  * 'fp32' (default of jmac_b200.finding; 3xTF32, three kind::tf32 MMAs per K step): compared with the oracle in
    FLOAT64 ON THE UN-ROUNDED INPUTS -- what the reference computes (f64 cdist / fp32 mm).  Tie band EPS32;
  * 'bf16' (the 1M x 1M mode): the oracle is fed the SAME bf16-rounded inputs; what remains is fp32
    accumulation-order noise.  Tie band EPS16.
Documented-tie rule (SURVEY 8c ii): indices / ranks must match exactly wherever the compared scores differ by more
than the band.  Small problems are cut stream-K style over all SMs; split and unsplit launches must agree bit for bit."""
import numpy as np
import pytest
import torch

from oracle import jmac_oracle as O

pytestmark = pytest.mark.gpu
EPS16 = 1e-5          # bf16 operands (products exact), fp32 accumulation order
EPS32 = 4e-6          # 3xTF32: ~2^-21 sum|a||b| (<= 1 for unit-norm rows) + fp32 accumulation in TMEM
BAND = {"bf16": EPS16, "fp32": EPS32}


def bf16_round(x: np.ndarray) -> np.ndarray:
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def unit(x):
    return x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)


def planted(n1, n2, d, seed, noise=1.0):
    rng = np.random.default_rng(seed)
    e1 = rng.standard_normal((n1, d)).astype(np.float32)
    e2 = rng.standard_normal((n2, d)).astype(np.float32)
    m = min(n1, n2)
    e2[:m] = e1[:m] + noise * rng.standard_normal((m, d)).astype(np.float32)
    return unit(e1).astype(np.float32), unit(e2).astype(np.float32)


def tables(F, e1, e2, prec, d):
    """Device tables of the kernel and the float64 matrices the oracle is entitled to for that precision."""
    t1, t2 = torch.from_numpy(e1).cuda(), torch.from_numpy(e2).cuda()
    if prec == "bf16":
        A, B = F.to_bf16_rows(t1), F.to_bf16_rows(t2)
        a, b = A.float().cpu().numpy()[:, :d], B.float().cpu().numpy()[:, :d]
        np.testing.assert_array_equal(a, bf16_round(e1))               # conversion kernel == round-to-nearest-even
        return A, B, a.astype(np.float64), b.astype(np.float64)
    A, B = F.to_tf32_rows(t1), F.to_tf32_rows(t2)
    hi, lo = A.hi.cpu().numpy()[:, :d], A.lo.cpu().numpy()[:, :d]
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)   # both terms are exact TF32 values
    assert float(np.abs(hi.astype(np.float64) + lo - e1).max()) <= 2.0 ** -21 * float(np.abs(e1).max())
    if A.hi.shape[1] > d:
        assert float(A.hi[:, d:].abs().max()) == 0.0 and float(A.lo[:, d:].abs().max()) == 0.0
    return A, B, e1.astype(np.float64), e2.astype(np.float64)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("n1,n2,d,k", [(300, 500, 64, 10), (1000, 777, 256, 10), (257, 129, 100, 5), (513, 2000, 256, 25), (64, 64, 32, 1),
                                         (700, 1500, 300, 25)])
def test_sim_topk_matches_float64(n1, n2, d, k, prec):
    from jmac_b200 import finding as F
    eps = BAND[prec]
    e1, e2 = planted(n1, n2, d, n1 + n2)
    A, B, a, b = tables(F, e1, e2, prec, d)
    S = a @ b.T
    gold = (np.arange(n1) * 7) % n2
    res = F.sim_topk(A, B, k, gold_col=torch.from_numpy(gold.astype(np.int32)).cuda(), want_diag=True)
    topv, topi, diag = res["topv"].cpu().numpy(), res["topi"].cpu().numpy(), res["diag"].cpu().numpy()
    np.testing.assert_allclose(diag, S[np.arange(n1), gold], atol=eps)
    ref_v = -np.sort(-S, axis=1)[:, :k]
    np.testing.assert_allclose(topv, ref_v, atol=eps)                  # values of the k largest, descending
    np.testing.assert_allclose(S[np.arange(n1)[:, None], topi], ref_v, atol=eps)   # indices point at them (ties: any of the equals)
    assert np.all(np.diff(topv, axis=1) <= 0)
    ref_i = np.argsort(-S, axis=1, kind="stable")[:, :k]
    gap_ok = np.abs(ref_v - -np.sort(-S, axis=1)[:, 1:k + 1]) > 2 * eps if n2 > k else np.ones_like(ref_v, dtype=bool)
    clear = np.concatenate([np.ones((n1, 1), bool), gap_ok[:, :-1]], 1) & gap_ok    # both neighbours further than the band
    assert np.array_equal(topi[clear], ref_i[clear])                   # bit-exact indices away from documented ties


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("n1,n2,d,k", [(2617, 2617, 256, 10), (300, 5000, 64, 25), (1500, 700, 128, 1), (4500, 30000, 256, 25), (256, 128, 64, 10)])
def test_split_launch_is_bit_identical_to_unsplit(n1, n2, d, k, prec):
    """Stream-K decomposition over all SMs + slot merge == one CTA per 256-row block sweeping every column."""
    from jmac_b200 import finding as F, _lib
    lib = _lib.load()
    e1, e2 = planted(n1, n2, d, 11 * n1 + n2)
    t1, t2 = torch.from_numpy(e1).cuda(), torch.from_numpy(e2).cuda()
    A, B = F.to_rows(t1, precision=prec), F.to_rows(t2, precision=prec)
    gold = torch.from_numpy(((np.arange(n1) * 13) % n2).astype(np.int32)).cuda()
    r1 = torch.rand(n1, device="cuda") * 0.1
    r2 = torch.rand(n2, device="cuda") * 0.1
    gv = torch.rand(n1, device="cuda") * 0.2
    if n2 > 128:
        assert lib.jmac_sim_grid(n1, n2, 1) > lib.jmac_sim_grid(n1, n2, 0) or lib.jmac_sim_grid(n1, n2, 0) >= lib.jmac_sm_count()
        assert lib.jmac_sim_grid(n1, n2, 1) <= lib.jmac_sm_count()
    for kw in (dict(), dict(alpha=2.0, r1=r1, r2=r2, gold_val=gv, want_count=True)):
        s = F.sim_topk(A, B, k, gold_col=gold, want_diag=True, split=True, **kw)
        u = F.sim_topk(A, B, k, gold_col=gold, want_diag=True, split=False, **kw)
        for key in ("topv", "topi", "diag", "count"):
            if s[key] is not None:
                assert torch.equal(s[key], u[key]), key
    mask1 = (torch.rand(n1, device="cuda") < 0.2).float()
    mask2 = (torch.rand(n2, device="cuda") < 0.2).float()
    s = F.softmax_stats(A, B, 20.0, row_mask=mask1, col_mask=mask2, split=True)
    u = F.softmax_stats(A, B, 20.0, row_mask=mask1, col_mask=mask2, split=False)
    assert torch.equal(s["max"], u["max"]) and torch.equal(s["arg"], u["arg"])
    torch.testing.assert_close(s["Z"], u["Z"], rtol=1e-4, atol=0)          # re-associated fp32 sums of up to 30 000 terms
    torch.testing.assert_close(s["W"], u["W"], rtol=2e-4, atol=1e-5)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("n,d,csls,metric", [(700, 256, 10, "cosine"), (515, 64, 5, "inner"), (300, 128, 0, "cosine"), (1000, 256, 0, "inner"),
                                              (4500, 256, 10, "cosine"),       # C3: the OpenEA-shape seed set (30 % of 15k)
                                              (900, 300, 10, "cosine"), (600, 300, 0, "inner")])   # JMAC_DBPv1 width
def test_alignment_ranks_match_oracle(n, d, csls, metric, prec):
    from jmac_b200 import finding as F
    eps = BAND[prec]
    e1, e2 = planted(n, n, d, n * 3 + csls, noise=1.5)
    if metric == "inner":          # un-normalised inputs for the inner-product metric
        e1, e2 = e1 * 1.7, e2 * 0.6
    ranks, top1 = F.alignment_ranks(e1, e2, metric=metric, normalize=False, csls_k=csls, precision=prec)
    ranks, top1 = ranks.cpu().numpy(), top1.cpu().numpy()
    rnd = bf16_round if prec == "bf16" else (lambda x: x)
    # bf16: the same fp32-normalised-then-rounded inputs; fp32: the reference's float64 cosine of the raw inputs
    a = rnd(unit(e1) if metric == "cosine" else e1).astype(np.float64)
    b = rnd(unit(e2) if metric == "cosine" else e2).astype(np.float64)
    if prec == "fp32" and metric == "cosine":
        a, b = unit(e1.astype(np.float64)), unit(e2.astype(np.float64))
    S = a @ b.T
    if csls > 0:
        r1 = -np.sort(-S, axis=1)[:, :csls].mean(1)
        r2 = -np.sort(-S.T, axis=1)[:, :csls].mean(1)
        S = 2 * S - r1[:, None] - r2[None, :]
    _, hits, mr, mrr, oranks, otop1 = O.greedy_alignment(None, None, (1, 5, 10), sim_mat=S)
    goldv = S[np.arange(n), np.arange(n)]
    near = (np.abs(S - goldv[:, None]) <= 4 * eps).sum(1) - 1          # competitors within the tie band
    assert np.all(np.abs(ranks - oranks) <= near)
    assert np.mean(ranks == oranks) > (0.999 if prec == "fp32" else 0.995)
    best = S.max(1)
    second = -np.sort(-S, axis=1)[:, 1]
    clear = (best - second) > 4 * eps
    assert np.array_equal(top1[clear], otop1[clear])


def test_greedy_alignment_api_and_reference_golden(golden):
    """API shape of evaluation.test and agreement with the golden metrics of the UNMODIFIED reference (float64 cosine,
    full argsort): in the default fp32 mode a Hits@k count may differ only by rows whose gold score has a competitor
    inside the fp32 band; with bf16 operands by the rows inside the (250x wider) bf16 band."""
    from jmac_b200 import finding as F
    g = golden("alignment")
    e1, e2 = g["e1"], g["e2"]
    n = e1.shape[0]
    S = {"cos_csls10": O.sim(e1.astype(np.float64), e2.astype(np.float64), "cosine", False, 10).astype(np.float64),
         "inner_nocsls": e1.astype(np.float64) @ e2.astype(np.float64).T}
    for name, kw in (("cos_csls10", dict(metric="cosine", csls_k=10)), ("inner_nocsls", dict(metric="inner", csls_k=0))):
        ref = g["test_" + name]
        goldv = S[name][np.arange(n), np.arange(n)]
        for prec, band in (("fp32", 2e-5), ("bf16", 2e-2)):
            top_k, hits, mr, mrr = F.greedy_alignment(e1, e2, [1, 5, 10], 8, normalize=False, accurate=True, logger=None, precision=prec, **kw)
            tied = int(((np.abs(S[name] - goldv[:, None]) <= band * np.abs(S[name]).max()).sum(1) > 1).sum())   # rows with a competitor in the band
            assert top_k == [1, 5, 10] and len(hits) == 3
            assert np.all(np.abs(np.asarray(hits) - ref[:3]) <= 100.0 * tied / n + 1e-9), (name, prec, hits, ref, tied)
            if tied == 0:
                assert abs(mr - ref[3]) <= 1e-9 and abs(mrr - ref[4]) <= 1e-9
            else:
                assert abs(mr - ref[3]) <= 0.15 and abs(mrr - ref[4]) <= 0.02
        # the public default is the fp32-accurate mode
        assert F.test(e1, e2, None, [1, 5, 10], 8, normalize=False, accurate=True, logger=None, **kw)[1].tolist() == \
            F.greedy_alignment(e1, e2, [1, 5, 10], 8, normalize=False, accurate=True, precision="fp32", **kw)[1].tolist()


def test_precomputed_matrix_and_other_metrics(golden):
    """alignment.py:40-41 (simi_matrix bypasses sim()) and the cdist-based metrics of similarity.py:41-52."""
    from jmac_b200 import finding as F
    g = golden("alignment")
    e1, e2 = g["e1"], g["e2"]
    n = e1.shape[0]
    ref = g["test_cos_csls10"]
    _, hits, mr, mrr = F.test(None, None, None, [1, 5, 10], 8, simi_matrix=g["sim_cos_csls10"], logger=None)
    np.testing.assert_allclose(list(hits) + [mr, mrr], ref, rtol=1e-9, atol=1e-9)
    a, b = e1.astype(np.float64), e2.astype(np.float64)
    for metric, dist in (("euclidean", np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))), ("manhattan", np.abs(a[:, None, :] - b[None, :, :]).sum(-1))):
        S = (1.0 - dist).astype(np.float32)
        np.testing.assert_allclose(F.sim(e1, e2, metric, False, 0), S, atol=2e-6)
        _, hits, mr, mrr = F.test(e1, e2, None, [1, 5, 10], 8, metric=metric, csls_k=0, logger=None)
        _, ohits, omr, omrr, _, _ = O.greedy_alignment(None, None, (1, 5, 10), sim_mat=S)
        assert np.all(np.abs(np.asarray(hits) - ohits) <= 100.0 / n + 1e-9) and abs(mrr - omrr) < 2e-3


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_get_neg_matches_oracle(prec):
    from jmac_b200 import finding as F
    eps = BAND[prec]
    e1, e2 = planted(400, 900, 256, 5)
    ill = np.random.default_rng(1).permutation(400)[:123]
    k = 25
    neg = F.get_neg(ill.tolist(), torch.from_numpy(e1), torch.from_numpy(e2), k, precision=prec)
    assert neg.dtype == torch.int64 and neg.shape == (123 * k,) and not neg.is_cuda
    rnd = bf16_round if prec == "bf16" else (lambda x: x)
    a, b = rnd(e1)[ill].astype(np.float64), rnd(e2).astype(np.float64)
    S = a @ b.T
    ref_v = -np.sort(-S, axis=1)[:, :k]
    got = neg.numpy().reshape(123, k)
    np.testing.assert_allclose(S[np.arange(123)[:, None], got], ref_v, atol=eps)
    oneg = O.get_neg(np.arange(123), torch.from_numpy(a), torch.from_numpy(b), k).numpy().reshape(123, k)
    kth_gap = ref_v[:, -1] - -np.sort(-S, axis=1)[:, k]                  # gap between the k-th and (k+1)-th neighbour
    clear = kth_gap > 2 * eps
    assert np.array_equal(np.sort(got[clear], 1), np.sort(oneg[clear], 1))       # identical index SETS away from documented ties
    assert np.mean(got == oneg) > 0.99
    if prec == "fp32":                                                    # the default is the fp32-accurate mode
        assert torch.equal(neg, F.get_neg(ill.tolist(), torch.from_numpy(e1), torch.from_numpy(e2), k))


def test_dense_sim_api(golden):
    from jmac_b200 import finding as F
    g = golden("alignment")
    np.testing.assert_allclose(F.sim(g["e1"], g["e2"], "cosine", False, 0), g["sim_cos"], atol=2e-6)
    np.testing.assert_allclose(F.sim(g["e1"], g["e2"], "cosine", False, 10), g["sim_cos_csls10"], atol=5e-6)
    np.testing.assert_allclose(F.csls_sim(g["sim_cos"], 10), g["sim_cos_csls10"], atol=5e-6)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("n1,n2,d,L", [(300, 260, 64, 90), (1000, 1100, 256, 400), (3000, 2600, 256, 2000)])
def test_compute_alignment_quality_matches_oracle(n1, n2, d, L, prec):
    """train.py:231-259 and its consumer (train.py:165-168) without the dense matrices."""
    from jmac_b200 import finding as F
    e1, e2 = planted(n1, n2, d, 77, noise=1.0)
    rng = np.random.default_rng(5)
    l1 = rng.permutation(min(n1, n2))[:L]
    l2 = l1.copy()                                  # aligned test pairs share the planted index
    t1, t2 = torch.from_numpy(e1), torch.from_numpy(e2)
    ent, simi, simi2 = F.compute_alignment_quality(t1, t2, l1.tolist(), l2.tolist(), precision=prec)
    rnd = bf16_round if prec == "bf16" else (lambda x: x)
    a, b = torch.from_numpy(rnd(e1)).double(), torch.from_numpy(rnd(e2)).double()
    oent, o1, o2 = O.compute_alignment_quality(a, b, l1, l2)
    rtol = 2e-4 if prec == "bf16" else 1e-4       # scale 20: a 4e-6 similarity error is 8e-5 in the logit
    np.testing.assert_allclose(float(ent), float(oent), rtol=rtol)
    assert not ent.is_cuda and ent.dim() == 0
    for view, dense in ((simi, o1), (simi2, o2)):
        mv, mi = view.max(dim=1)
        ov, oi = dense.max(1)
        np.testing.assert_allclose(mv.numpy(), ov.numpy(), rtol=rtol, atol=1e-7)
        sd = torch.sort(dense, dim=1, descending=True)[0]
        clear = ((sd[:, 0] - sd[:, 1]) > 1e-4 * sd[:, 0]).numpy()          # documented ties (incl. fully masked rows: uniform)
        assert np.array_equal(mi.numpy()[clear], oi.numpy()[clear])
        assert len(view) == dense.shape[0] and view.shape == tuple(dense.shape)
    # the reference's consumer code runs unchanged on the view (train.py:165-168)
    torch.manual_seed(0)
    max_values = simi.max(dim=1)[0]
    src_nodes = max_values.multinomial(20, replacement=False).tolist()
    rows = simi[src_nodes]
    dst_nodes = rows.max(dim=1)[1].tolist()
    assert len(dst_nodes) == 20 and all(0 <= x < n2 for x in dst_nodes)
    assert dst_nodes == simi.max(dim=1)[1][src_nodes].tolist()


def test_wide_rows_fp32_mode():
    """d > 320 (no resident A block needed in the 3xTF32 mode: both operands stream through the ring)."""
    from jmac_b200 import finding as F
    e1, e2 = planted(500, 640, 512, 3)
    ranks, _ = F.alignment_ranks(e1, e2, metric="cosine", csls_k=0)
    S = unit(e1.astype(np.float64)) @ unit(e2.astype(np.float64)).T
    oranks = (S > S[np.arange(500), np.arange(500)][:, None]).sum(1)
    near = (np.abs(S - S[np.arange(500), np.arange(500)][:, None]) <= 4 * EPS32).sum(1) - 1
    assert np.all(np.abs(ranks.cpu().numpy() - oranks) <= near)
    with pytest.raises(NotImplementedError):
        F.alignment_ranks(e1, e2, metric="cosine", csls_k=0, precision="bf16")
