"""GPU parity of the completion scorer, pair cosine, link prediction kernels and of the JMAC
drop-in's losses / evaluator against the oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import jmac_oracle as O

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


def make_args(d, **kw):
    from types import SimpleNamespace
    a = SimpleNamespace(dim=d, device="cuda", dropout=0.0, leaky_relu_w=0.05, comp_op="sub", margin_align=1.0, num_negative=3,
                        margin_completion=5.0, num_gcn_layer=2, no_name_info=True, batch_size=8)
    a.__dict__.update(kw)
    return a


@pytest.mark.parametrize("N,R,d,B,K,tiled", [(500, 17, 256, 64, 25, "yes"), (300, 5, 300, 33, 7, "yes"), (200, 4, 64, 16, 70, "no"),
                                              (100, 3, 32, 1, 1, "yes"),
                                              (40, 3, 64, 200, 30, "yes"),        # few rows: every row's occurrences span many reduction chunks
                                              (60, 2, 128, 150, 40, "partly"),    # leading (h, r) runs shorter than K+1: dense record + mask occurrences
                                              (3000, 11, 64, 4096, 127, "yes"),   # 2M + B >= 2^20 occurrences: 128-wide chunks
                                              (700, 9, 512, 40, 12, "no"), (256, 5, 400, 31, 9, "partly")])
def test_transe_margin_matches_oracle(N, R, d, B, K, tiled):
    from jmac_b200 import functional as JF
    rng = np.random.default_rng(B * 131 + K)
    torch.manual_seed(B)
    ents = [torch.randn(N, d) for _ in range(2)]
    rels = [torch.randn(R, d) for _ in range(2)]
    ents[1][:, : d // 4] = 0.0                    # exact zeros in h + r - t: sign(0) = 0, as torch.abs' gradient
    rels[1][:, : d // 4] = 0.0
    M = (K + 1) * B
    if tiled == "yes":   # train.py:347-352
        h = np.tile(rng.integers(0, N, B), K + 1)
        r = np.tile(rng.integers(0, R, B), K + 1)
    elif tiled == "no":  # arbitrary batch_h / batch_r (general contract of completion_loss)
        h, r = rng.integers(0, N, M), rng.integers(0, R, M)
    else:                # tiled up to a per-positive point, arbitrary behind it; a third of the positives fully tiled
        h = np.tile(rng.integers(0, N, B), K + 1).reshape(K + 1, B)
        r = np.tile(rng.integers(0, R, B), K + 1).reshape(K + 1, B)
        cut = rng.integers(1, K + 2, B)
        cut[::3] = K + 1
        for b in range(B):
            h[cut[b]:, b] = rng.integers(0, N, K + 1 - cut[b])
            r[cut[b]:, b] = rng.integers(0, R, K + 1 - cut[b])
        h, r = h.reshape(-1), r.reshape(-1)
    t = rng.integers(0, N, M)
    margin = 5.0 if d <= 300 else 40.0            # keep a mix of active and inactive hinge terms at every width
    ge = [e.cuda().requires_grad_(True) for e in ents]
    gr = [x.cuda().requires_grad_(True) for x in rels]
    loss = JF.transe_margin_loss(ge, gr, T(h), T(r), T(t), B, margin)
    w = torch.tensor([0.7, 1.3], device="cuda")
    (loss * w).sum().backward()
    oe = [e.double().requires_grad_(True) for e in ents]
    orr = [x.double().requires_grad_(True) for x in rels]
    ol = torch.stack([O.transe_margin(oe[l], orr[l], T(h), T(r), T(t), B, margin)[0].reshape(()) for l in range(2)])
    (ol * w.cpu().double()).sum().backward()
    np.testing.assert_allclose(loss.detach().cpu().numpy(), ol.detach().numpy(), rtol=1e-5)
    for l in range(2):
        np.testing.assert_allclose(ge[l].grad.cpu().numpy(), oe[l].grad.numpy(), atol=1e-6, rtol=1e-4)
        np.testing.assert_allclose(gr[l].grad.cpu().numpy(), orr[l].grad.numpy(), atol=1e-6, rtol=1e-4)
    # atomics-free backward: bit-identical from run to run
    first = [g_.grad.clone() for g_ in ge + gr]
    for g_ in ge + gr:
        g_.grad = None
    (JF.transe_margin_loss(ge, gr, T(h), T(r), T(t), B, margin) * w).sum().backward()
    for a, b_ in zip(first, [g_.grad for g_ in ge + gr]):
        assert torch.equal(a, b_)
    # no gradient requested: same loss through the record-free forward
    with torch.no_grad():
        assert torch.equal(JF.transe_margin_loss(ge, gr, T(h), T(r), T(t), B, margin), loss.detach())


def test_transe_bad_layout_raises_like_reference():
    from jmac_b200 import functional as JF
    e = [torch.randn(10, 32).cuda()]
    r = [torch.randn(3, 32).cuda()]
    idx = torch.zeros(14, dtype=torch.long)
    with pytest.raises(RuntimeError):          # 14 triples cannot be viewed as [-1, 4] after the 4 positives
        JF.transe_margin_loss(e, r, idx, idx, idx, 4, 5.0)


@pytest.mark.parametrize("d", [32, 256, 300])
def test_pair_cos_matches_oracle(d):
    from jmac_b200 import functional as JF
    rng = np.random.default_rng(d)
    torch.manual_seed(d)
    A, B = torch.randn(200, d), torch.randn(150, d)
    A[3] = 0                                             # zero row: F.normalize clamps at 1e-12
    ia, ib = rng.integers(0, 200, 400), rng.integers(0, 150, 400)
    ia[0] = 3
    g = torch.randn(400)
    Ag, Bg = A.cuda().requires_grad_(True), B.cuda().requires_grad_(True)
    cosv = JF.pair_cos(Ag, Bg, ia.astype(np.float64), T(ib))          # float64 indices as train.py:189-192 produces
    (cosv * g.cuda()).sum().backward()
    Ao, Bo = A.double().requires_grad_(True), B.double().requires_grad_(True)
    co = (torch.nn.functional.normalize(Ao[T(ia)], 2, -1) * torch.nn.functional.normalize(Bo[T(ib)], 2, -1)).sum(1)
    (co * g.double()).sum().backward()
    np.testing.assert_allclose(cosv.detach().cpu().numpy(), co.detach().numpy(), atol=1e-6, rtol=1e-5)
    np.testing.assert_allclose(Ag.grad.cpu().numpy(), Ao.grad.numpy(), atol=2e-6, rtol=1e-4)
    np.testing.assert_allclose(Bg.grad.cpu().numpy(), Bo.grad.numpy(), atol=2e-6, rtol=1e-4)


@pytest.mark.parametrize("N,d,B", [(1000, 256, 130), (333, 300, 7), (129, 32, 257)])
def test_l1_dist_and_rank_match_oracle(N, d, B):
    from jmac_b200 import functional as JF
    rng = np.random.default_rng(N)
    torch.manual_seed(N)
    R = 6
    ents = [torch.randn(N, d), torch.randn(N, d)]
    rels = [torch.randn(R, d), torch.randn(R, d)]
    e, r, t = rng.integers(0, N, B), rng.integers(0, R, B), rng.integers(0, N, B)
    ce, cr = [x.cuda() for x in ents], [x.cuda() for x in rels]
    for pred_head in (False, True):
        dist = JF.l1_dist(ce, cr, e.tolist(), r.tolist(), pred_head)
        od = O.forward_linkpred(ents, rels, e, r, pred_head)
        np.testing.assert_allclose(dist.cpu().numpy(), od.numpy(), rtol=2e-6, atol=1e-4)
    # filtered ranks: CSR filter lists (true tails), the target is always among them
    true_tail = {}
    for j in range(B):
        extra = rng.integers(0, N, rng.integers(0, 6))
        true_tail[(int(e[j]), int(r[j]))] = np.unique(np.concatenate([true_tail.get((int(e[j]), int(r[j])), np.zeros(0, np.int64)), extra, [t[j]]]).astype(np.int64))
    lists = [true_tail[(int(e[j]), int(r[j]))] for j in range(B)]
    fptr = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    fidx = np.concatenate(lists)
    # (1) against the reference's sort-based ranking applied to the kernel's own fp32 distances:
    #     bit-exact (rank-by-counting == position in the descending sort when there are no exact ties)
    pred32 = -JF.l1_dist(ce, cr, e, r).cpu()
    # (2) against the fp64 oracle: identical except for documented ties, i.e. competitors whose
    #     distance is within 1e-5 relative of the target's (fp32 rounding can order them either way)
    pred64 = -O.forward_linkpred([x.double() for x in ents], [x.double() for x in rels], e, r)
    for filt in (True, False):
        ranks = JF.l1_rank(ce, cr, e, r, t, fptr if filt else None, fidx if filt else None).cpu().numpy()
        assert np.array_equal(ranks, O.linkpred_ranks(pred32, e, r, t, true_tail if filt else None))
        oranks = O.linkpred_ranks(pred64, e, r, t, true_tail if filt else None)
        tgt = pred64[torch.arange(B), T(t)]
        near = ((pred64 - tgt[:, None]).abs() <= 1e-5 * tgt.abs()[:, None]).sum(1).numpy() - 1
        assert np.all(np.abs(ranks - oranks) <= near)


def model_pair(g, d_model_args):
    from jmac_b200 import JMAC
    sd = {k[3:]: T(v) for k, v in g.items() if k.startswith("sd.")}
    n_ent = sd["ent_init_att_completion"].shape[0]
    model = JMAC(d_model_args, np.zeros((n_ent, 300), dtype=np.float32), sd["rel_init_att_completion"].shape[0], n_ent).cuda()
    missing = model.load_state_dict(sd, strict=True)
    return model


def test_state_dict_keys_match_reference(golden):
    g = golden("model_noname")
    ref_keys = sorted(k[3:] for k in g if k.startswith("sd."))
    model = model_pair(g, make_args(16))
    assert sorted(model.state_dict().keys()) == ref_keys


@pytest.mark.parametrize("source", [True, False])
def test_completion_loss_matches_reference_golden(golden, source):
    g = golden("model_noname")
    n1, n2, nr, B, K = (int(g[k]) for k in ("n1", "n2", "nr", "B", "K"))
    model = model_pair(g, make_args(16, num_negative=K, batch_size=B))
    model.train()
    fd = {"links": g["links"], "ent_bases1": [0, n1], "ent_bases2": [n1, n1 + n2], "rel_bases1": [0, nr], "rel_bases2": [nr, 2 * nr]}
    h, r = np.tile(g["sub"], K + 1), np.tile(g["rel"], K + 1)
    t = np.concatenate([g["obj"], g["neg"].reshape(-1)])
    if not source:
        h, t = h % n2, t % n2
    data = {"batch_h": T(h).cuda(), "batch_r": T(r).cuda(), "batch_t": T(t).cuda()}
    loss = model.completion_loss(data, T(g["ei1"]).cuda(), T(g["et1"]).cuda(), T(g["ei2"]).cuda(), T(g["et2"]).cuda(), fd, source)
    tag = "src" if source else "dst"
    np.testing.assert_allclose(loss.item(), float(g["closs_" + tag].reshape(-1)[0]), rtol=1e-4)
    loss.backward()
    for k, v in g.items():
        if k.startswith(f"cgrad_{tag}."):
            name = k.split(".", 1)[1]
            p = dict(model.named_parameters())[name]
            got = np.zeros_like(v) if p.grad is None else p.grad.cpu().numpy()
            scale = max(float(np.abs(v).max()), 1e-3)
            assert float(np.abs(got - v).max()) <= 5e-3 * scale + 2e-5, name


def test_alignment_loss_linkpred_getemb_match_reference_golden(golden):
    g = golden("model_noname")
    n1, n2, nr, B, K = (int(g[k]) for k in ("n1", "n2", "nr", "B", "K"))
    model = model_pair(g, make_args(16, num_negative=K, batch_size=B))
    model.train()
    fa = {"links": g["links"], "ent_bases1": [0, n1], "ent_bases2": [n1, n1 + n2], "rel_bases1": [0, nr], "rel_bases2": [nr, 2 * nr],
          "neg_left": g["fa.neg_left"], "neg_right": T(g["fa.neg_right"]), "neg2_left": T(g["fa.neg2_left"]), "neg2_right": g["fa.neg2_right"]}
    ei1, et1, ei2, et2 = (T(g[k]).cuda() for k in ("ei1", "et1", "ei2", "et2"))
    loss = model.alignment_loss(fa, ei1, et1, ei2, et2)
    np.testing.assert_allclose(loss.item(), float(g["aloss"]), rtol=1e-4)
    loss.backward()
    for k, v in g.items():
        if k.startswith("agrad."):
            name = k.split(".", 1)[1]
            p = dict(model.named_parameters())[name]
            got = np.zeros_like(v) if p.grad is None else p.grad.cpu().numpy()
            scale = max(float(np.abs(v).max()), 1e-3)
            assert float(np.abs(got - v).max()) <= 5e-3 * scale + 2e-5, name
    # empty links: python 0, as the reference (src/jmac_model.py:241-242,257-258)
    assert model.alignment_loss({**fa, "links": []}, ei1, et1, ei2, et2) == 0
    assert model.alignment_loss_simple([], None, None) == 0
    # eval mode: link prediction + get_emb
    sd = {k[3:]: T(v) for k, v in g.items() if k.startswith("sd.")}
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        dist = model.forward_linkpred(g["lp_e"].tolist(), g["lp_r"].tolist(), ei1, et1, list(range(n1)), [0, n1], [0, nr])
        dh = model.forward_linkpred(g["lp_e"].tolist(), g["lp_r"].tolist(), ei1, et1, list(range(n1)), [0, n1], [0, nr], pred_head=True)
        ea, ec = model.get_emb(ei1, et1, [0, n1], [0, nr])
    np.testing.assert_allclose(dist.cpu().numpy(), g["lp_dist"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dh.cpu().numpy(), g["lp_dist_head"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ec, g["emb_comp"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(ea, g["emb_align"], atol=1e-5, rtol=1e-4)


def test_completion_evaluator_matches_reference_golden(golden):
    """Same metrics as the reference's CompletionEvaluator.test on a small KG, val/test x raw/filtered."""
    from types import SimpleNamespace
    from jmac_b200 import CompletionEvaluator
    g = golden("evaluator")
    n, nr = int(g["n"]), int(g["nr"])
    model = model_pair(g, make_args(16, batch_size=16))
    model.eval()
    allt = np.concatenate([g["train"], g["val"], g["test"]])
    true_tail = {}
    for h, r, t in allt:
        true_tail.setdefault((int(h), int(r)), set()).add(int(t))

    def gen(h, r, t, batch_size, shuffle=False):
        tri = torch.stack([h, r, t], 1)
        return [tri[i:i + batch_size] for i in range(0, len(tri), batch_size)]

    kg = SimpleNamespace(edge_index=g["ei"], edge_type=g["et"], entity_id_base=0, upper_entity_base=n, relation_id_base=0,
                         upper_relation_base=nr, true_tail={k: np.array(sorted(v)) for k, v in true_tail.items()}, num_entity=n,
                         h_val=T(g["val"][:, 0]), r_val=T(g["val"][:, 1]), t_val=T(g["val"][:, 2]),
                         h_test=T(g["test"][:, 0]), r_test=T(g["test"][:, 1]), t_test=T(g["test"][:, 2]), generate_batch_data=gen)
    ev = CompletionEvaluator(kg, model, "cuda", None)
    args = make_args(16, batch_size=16)
    with torch.no_grad():
        for is_val in (1, 0):
            for filt in (1, 0):
                got = ev.test(args, is_val=bool(is_val), filterr=bool(filt))
                np.testing.assert_allclose(got, g[f"metrics_val{is_val}_filt{filt}"], rtol=1e-6)


def test_dbpv1_flags_normalized_triples_and_relu_relation_transform(golden):
    """The JMAC_DBPv1 variant's differences on this path are flags of the same kernels: unit-norm h, r, t
    in the completion scorer (JMAC_DBPv1/models/jmac_model.py:245-250) and ReLU in the relation transform (:27,:51).
    Checked against the oracle fed the same state dict."""
    from oracle import jmac_oracle as O
    g = golden("model_noname")
    n1, n2, nr, B, K = (int(g[k]) for k in ("n1", "n2", "nr", "B", "K"))
    model = model_pair(g, make_args(16, num_negative=K, batch_size=B, normalize_triples=True))
    model.train()
    fd = {"links": g["links"], "ent_bases1": [0, n1], "ent_bases2": [n1, n1 + n2], "rel_bases1": [0, nr], "rel_bases2": [nr, 2 * nr]}
    h, r = np.tile(g["sub"], K + 1), np.tile(g["rel"], K + 1)
    t = np.concatenate([g["obj"], g["neg"].reshape(-1)])
    data = {"batch_h": T(h).cuda(), "batch_r": T(r).cuda(), "batch_t": T(t).cuda()}
    loss = model.completion_loss(data, T(g["ei1"]).cuda(), T(g["et1"]).cuda(), T(g["ei2"]).cuda(), T(g["et2"]).cuda(), fd, True)
    sd = {k[3:]: T(v).double() for k, v in g.items() if k.startswith("sd.")}
    _, e1, r1 = O.forward_no_name(sd, T(g["ei1"]), T(g["et1"]), [0, n1], [0, nr])
    _, e2, _ = O.forward_no_name(sd, T(g["ei2"]), T(g["et2"]), [n1, n1 + n2], [nr, 2 * nr])
    nz = lambda x: torch.nn.functional.normalize(x, 2, -1)
    want = 0
    for layer in range(2):
        l_res, _ = O.transe_margin(nz(e1[layer]), nz(r1[layer]), T(h), T(r), T(t), B, 5.0)
        want = want + l_res + O.alignment_loss_simple(g["links"], e1[layer], e2[layer])
    np.testing.assert_allclose(loss.item(), float(want), rtol=1e-4)
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    # ReLU relation transform
    layer = model.conv1_completion
    layer.rel_act = "relu"
    rel = torch.randn(5, 16, device="cuda")
    got = layer.relation_transform(rel)
    want = torch.relu(torch.cat([rel, layer.loop_rel]) @ layer.rel_transform_weight1) @ layer.rel_transform_weight2
    torch.testing.assert_close(got, want)


def test_name_encoder_matches_reference_golden(golden):
    """JMAC.forward_name (src/jmac_model.py:172-204): name projection, three RelationAwareLayers, the concat-linear
    fusions; forward, both losses and every parameter gradient against the unmodified reference."""
    from jmac_b200 import JMAC
    g = golden("model_name")
    n1, n2, nr, B, K = (int(g[k]) for k in ("n1", "n2", "nr", "B", "K"))
    args = make_args(32, num_negative=K, batch_size=B, no_name_info=False)
    model = JMAC(args, g["name_emb"], 2 * nr, n1 + n2).cuda()
    sd = {k[3:]: T(v) for k, v in g.items() if k.startswith("sd.")}
    model.load_state_dict(sd)
    model.train()
    ei1, et1, ei2, et2 = (T(g[k]).cuda() for k in ("ei1", "et1", "ei2", "et2"))
    out, cembs, crels = model.forward_name(ei1, et1, [0, n1], [0, nr])
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["fwd_out"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(cembs[1].detach().cpu().numpy(), g["fwd_comp1"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(crels[1].detach().cpu().numpy(), g["fwd_rel1"], atol=1e-5, rtol=1e-4)
    fd = {"links": g["links"], "ent_bases1": [0, n1], "ent_bases2": [n1, n1 + n2], "rel_bases1": [0, nr], "rel_bases2": [nr, 2 * nr]}

    def check_grads(prefix):
        n = 0
        for k, v in g.items():
            if k.startswith(prefix):
                p = dict(model.named_parameters())[k.split(".", 1)[1]]
                got = np.zeros_like(v) if p.grad is None else p.grad.cpu().numpy()
                scale = max(float(np.abs(v).max()), 1e-3)
                assert float(np.abs(got - v).max()) <= 5e-3 * scale + 2e-5, k
                n += 1
        return n

    data = {"batch_h": T(np.tile(g["sub"], K + 1)).cuda(), "batch_r": T(np.tile(g["rel"], K + 1)).cuda(),
            "batch_t": T(np.concatenate([g["obj"], g["neg"].reshape(-1)])).cuda()}
    model.zero_grad()
    closs = model.completion_loss(data, ei1, et1, ei2, et2, fd, True)
    np.testing.assert_allclose(closs.item(), float(g["closs"].reshape(-1)[0]), rtol=1e-4)
    closs.backward()
    assert check_grads("cgrad.") >= 10
    model.zero_grad()
    fa = {**fd, "neg_left": g["fa.neg_left"], "neg_right": T(g["fa.neg_right"]), "neg2_left": T(g["fa.neg2_left"]), "neg2_right": g["fa.neg2_right"]}
    aloss = model.alignment_loss(fa, ei1, et1, ei2, et2)
    np.testing.assert_allclose(aloss.item(), float(g["aloss"]), rtol=1e-4)
    aloss.backward()
    assert check_grads("agrad.") >= 30
    model.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        ea, ec = model.get_emb(ei1, et1, [0, n1], [0, nr])
    np.testing.assert_allclose(ea, g["emb_align"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(ec, g["emb_comp"], atol=1e-5, rtol=1e-4)


def test_completion_loss_multi_equals_separate_calls(golden):
    """Optional multi-batch step: one encoder pass, several batches -- every loss equals completion_loss on that batch,
    and the accumulated gradient equals the sum of the separate gradients."""
    g = golden("model_noname")
    n1, n2, nr, B, K = (int(g[k]) for k in ("n1", "n2", "nr", "B", "K"))
    model = model_pair(g, make_args(16, num_negative=K, batch_size=B))
    model.train()
    fd = {"links": g["links"], "ent_bases1": [0, n1], "ent_bases2": [n1, n1 + n2], "rel_bases1": [0, nr], "rel_bases2": [nr, 2 * nr]}
    graphs = [T(g[k]).cuda() for k in ("ei1", "et1", "ei2", "et2")]
    rng = np.random.default_rng(3)
    batches = []
    for _ in range(3):
        h, r = np.tile(rng.integers(0, n1, B), K + 1), np.tile(rng.integers(0, nr, B), K + 1)
        batches.append({"batch_h": T(h).cuda(), "batch_r": T(r).cuda(), "batch_t": T(rng.integers(0, n1, (K + 1) * B)).cuda()})
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    want_loss, want_grad = [], None
    for b in batches:
        model.load_state_dict(sd0)          # BatchNorm running statistics back to the start
        model.zero_grad()
        l = model.completion_loss(b, *graphs, fd, True)
        l.backward()
        want_loss.append(l.item())
        gr = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        want_grad = gr if want_grad is None else {k: want_grad[k] + v for k, v in gr.items()}
    model.load_state_dict(sd0)
    model.zero_grad()
    losses = model.completion_loss_multi(batches, *graphs, fd, True)
    np.testing.assert_allclose([l.item() for l in losses], want_loss, rtol=1e-6)
    sum(losses).backward()
    for k, p in model.named_parameters():
        if k in want_grad:
            scale = max(float(want_grad[k].abs().max()), 1e-6)
            atol = 2e-6 if k.endswith("loop_rel") else 1e-7        # loop_rel: mathematically zero in train mode, rounding noise only
            assert float((p.grad - want_grad[k]).abs().max()) <= 1e-4 * scale + atol, k


def test_negative_sampler_on_device_feeds_the_scorer():
    """jmac_b200.sampling on CUDA tensors: valid, distinct negatives in the layout completion_loss expects (train.py:347-352)."""
    from jmac_b200 import functional as JF
    from jmac_b200.sampling import TripleIndex, sample_negatives
    rng = np.random.default_rng(5)
    n_ent, n_rel, B, K, d = 3000, 11, 512, 25, 64
    tri = np.stack([rng.integers(0, n_ent, 20000), rng.integers(0, n_rel, 20000), rng.integers(0, n_ent, 20000)], 1)
    idx = TripleIndex(tri, n_ent, n_rel, device="cuda")
    batch = torch.from_numpy(tri[:B]).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    neg = sample_negatives(idx, batch[:, 0], batch[:, 1], K, generator=g)
    assert neg.is_cuda and neg.shape == (B, K)
    assert not bool(idx.contains(batch[:, :1], batch[:, 1:2], neg).any())
    assert bool((torch.sort(neg, 1)[0][:, 1:] != torch.sort(neg, 1)[0][:, :-1]).all())
    ents = [torch.randn(n_ent, d, device="cuda", requires_grad=True)]
    rels = [torch.randn(n_rel, d, device="cuda", requires_grad=True)]
    loss = JF.transe_margin_loss(ents, rels, batch[:, 0].repeat(K + 1), batch[:, 1].repeat(K + 1),
                                 torch.cat((batch[:, 2], neg.t().reshape(-1))), B, 5.0)      # negatives [K, B] flattened: position B + k*B + b
    loss.sum().backward()
    assert torch.isfinite(loss).all() and ents[0].grad is not None
    # reproducible under a seeded generator, different under another seed
    again = sample_negatives(idx, batch[:, 0], batch[:, 1], K, generator=torch.Generator(device="cuda").manual_seed(1))
    other = sample_negatives(idx, batch[:, 0], batch[:, 1], K, generator=torch.Generator(device="cuda").manual_seed(2))
    assert torch.equal(neg, again) and not torch.equal(neg, other)


def test_negative_sampler_kernel_is_uniform_and_exhaustive():
    """The kernel's distribution: (a) with exactly k valid entities every row returns all of them; (b) chi-square test of
    uniformity over the valid entities of one (h, r), against the same statistic of the torch reference procedure."""
    import jmac_b200
    from jmac_b200.sampling import TripleIndex, sample_negatives, sample_negatives_torch
    n_ent, n_rel = 12, 2
    tri = np.array([[0, 0, t] for t in range(9)] + [[1, 0, 5]])                       # (0,0) has 9 known tails: valid = {9, 10, 11}
    idx = TripleIndex(tri, n_ent, n_rel, device="cuda")
    neg = sample_negatives(idx, torch.tensor([0, 1, 0]), torch.tensor([0, 0, 0]), 3)
    assert sorted(neg[0].tolist()) == [9, 10, 11] and sorted(neg[2].tolist()) == [9, 10, 11] and 5 not in neg[1].tolist()
    with pytest.raises(ValueError):
        sample_negatives(idx, torch.tensor([0]), torch.tensor([0]), 4)
    assert sample_negatives(idx, torch.tensor([], dtype=torch.long), torch.tensor([], dtype=torch.long), 3).shape == (0, 3)
    jmac_b200.check_errors()
    # uniformity: 50 entities, 10 excluded, k = 4, 40000 rows -> every valid entity appears with probability k / 40
    n_ent, k, B = 50, 4, 40000
    tri = np.array([[0, 0, t] for t in range(10)])
    idx = TripleIndex(tri, n_ent, 1, device="cuda")
    z = torch.zeros(B, dtype=torch.long)
    torch.manual_seed(7)
    for fn in (sample_negatives, sample_negatives_torch):
        neg = fn(idx, z, z, k)
        cnt = torch.bincount(neg.reshape(-1).cpu(), minlength=n_ent).double()
        assert float(cnt[:10].sum()) == 0
        exp = B * k / 40.0
        chi2 = float(((cnt[10:] - exp) ** 2 / exp).sum())
        assert chi2 < 80.0, (fn.__name__, chi2)                                        # 39 degrees of freedom: P(chi2 > 80) ~ 1e-4
        first = torch.bincount(neg[:, 0].cpu(), minlength=n_ent).double()[10:]        # the first pick alone is uniform too
        assert float(((first - B / 40.0) ** 2 / (B / 40.0)).sum()) < 80.0


def test_out_of_range_indices_are_flagged_not_followed():
    """The reference raises IndexError (CPU) / a device-side assert (CUDA) on an index outside its table.  The kernels here
    substitute row 0 and raise a flag in mapped host memory: nothing is read or written out of bounds (in particular the
    gradient REDs of the scorers' backward), and the error surfaces as IndexError at the next index-consuming call or at
    jmac_b200.check_errors()."""
    import jmac_b200
    from jmac_b200 import functional as JF
    N, R, d, B, K = 500, 7, 64, 16, 4
    torch.manual_seed(0)
    ent = torch.randn(N, d, device="cuda", requires_grad=True)
    rel = torch.randn(R, d, device="cuda", requires_grad=True)
    guard = torch.zeros(4096, device="cuda")                        # allocated right behind the tables: must stay untouched
    rng = np.random.default_rng(0)
    h, r = np.tile(rng.integers(0, N, B), K + 1), np.tile(rng.integers(0, R, B), K + 1)
    t = rng.integers(0, N, (K + 1) * B)
    jmac_b200.check_errors()                                         # clean slate
    loss = JF.transe_margin_loss([ent], [rel], h, r, t, B, 5.0)
    loss.sum().backward()
    jmac_b200.check_errors()                                         # valid indices: no flag
    for bad_arr, name in ((t, "t"), (h, "h"), (r, "r")):
        bad = bad_arr.copy()
        bad[3] = 10 ** 6 if name != "r" else R                       # far outside / one past the end
        args = {"h": h, "r": r, "t": t}
        args[name] = bad
        ent.grad = rel.grad = None
        loss = JF.transe_margin_loss([ent], [rel], args["h"], args["r"], args["t"], B, 5.0)
        loss.sum().backward()
        with pytest.raises(IndexError):
            jmac_b200.check_errors()
        assert torch.isfinite(ent.grad).all() and float(guard.abs().max()) == 0.0
        jmac_b200.check_errors()                                     # the flag was consumed
    # negative index
    bad = t.copy()
    bad[0] = -1
    JF.transe_margin_loss([ent.detach()], [rel.detach()], h, r, bad, B, 5.0)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):                                  # surfaces at the next index-consuming call without an explicit check
        JF.transe_margin_loss([ent.detach()], [rel.detach()], h, r, t, B, 5.0)
    # pair cosine and link prediction
    A, Bm = torch.randn(50, d, device="cuda", requires_grad=True), torch.randn(60, d, device="cuda", requires_grad=True)
    JF.pair_cos(A, Bm, [1, 2, 50], [3, 4, 5]).sum().backward()
    with pytest.raises(IndexError):
        jmac_b200.check_errors()
    JF.l1_rank([ent.detach()], [rel.detach()], [1, 2], [0, R + 3], [5, 6])
    with pytest.raises(IndexError):
        jmac_b200.check_errors()
    JF.l1_dist([ent.detach()], [rel.detach()], [1, N], [0, 1])
    with pytest.raises(IndexError):
        jmac_b200.check_errors()
    with pytest.raises(ValueError):                                  # one index list addresses every layer's table: shapes must agree
        JF.transe_margin_loss([ent.detach(), ent.detach()[:100]], [rel.detach(), rel.detach()], h, r, t, B, 5.0)
    jmac_b200.check_errors()


def test_forward_linkpred_index_list_and_gradient(golden):
    """forward_linkpred with an all_index that is NOT range(N) selects / permutes the table first (src/jmac_model.py:304-306),
    and the distances carry a gradient when one is requested."""
    g = golden("model_noname")
    n1, nr = int(g["n1"]), int(g["nr"])
    model = model_pair(g, make_args(16))
    model.eval()
    ei1, et1 = T(g["ei1"]).cuda(), T(g["et1"]).cuda()
    perm = np.random.default_rng(0).permutation(n1)
    e, r = [0, 5, 7], [1, 2, 3]
    with torch.no_grad():
        full = model.forward_linkpred(e, r, ei1, et1, list(range(n1)), [0, n1], [0, nr])
        sub = model.forward_linkpred(e, r, ei1, et1, perm.tolist(), [0, n1], [0, nr])
        # reference semantics: table' = table[perm]; the query rows e index table'
        ref = model.forward_linkpred(perm[e].tolist(), r, ei1, et1, list(range(n1)), [0, n1], [0, nr])[:, perm]
    torch.testing.assert_close(sub, ref, rtol=1e-5, atol=1e-4)
    assert not torch.equal(sub, full)
    # a permutation that keeps the first and the last entry in place is still a permutation
    p2 = np.arange(n1)
    p2[[3, 4]] = p2[[4, 3]]
    with torch.no_grad():
        swapped = model.forward_linkpred(e, r, ei1, et1, p2.tolist(), [0, n1], [0, nr])
    torch.testing.assert_close(swapped[:, 3], full[:, 4], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(swapped[:, 10], full[:, 10], rtol=1e-5, atol=1e-4)
    # gradient path (differentiable library cdist): same values, and a gradient arrives at the embedding table
    model.train()
    dist = model.forward_linkpred(e, r, ei1, et1, list(range(n1)), [0, n1], [0, nr])
    assert dist.requires_grad
    dist.sum().backward()
    assert model.ent_init_att_completion.grad is not None and float(model.ent_init_att_completion.grad.abs().sum()) > 0
    # the frozen margin is re-read after a load_state_dict
    m0 = model._margin()
    sd2 = {k: v.clone() for k, v in model.state_dict().items()}
    sd2["margin_completion"] = sd2["margin_completion"] + 2.0
    model.load_state_dict(sd2)
    assert abs(model._margin() - (m0 + 2.0)) < 1e-6
