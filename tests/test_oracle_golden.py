"""The oracle (oracle/jmac_oracle.py) against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  CPU only.  Tolerances: fp32 forward atol 1e-5 + rtol 1e-4 (north_star);
gradients atol 2e-5 + rtol 1e-3 in fp32 (the reference's own fp32 noise)."""
import numpy as np
import pytest
import torch

from oracle import jmac_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def layer_from_golden(g, dtype=torch.float32):
    p = {k: T(g["sd." + k]).to(dtype).clone().requires_grad_(True) for k in O.LAYER_PARAM_NAMES}
    st = O.BNState(T(g["sd.bn.running_mean"]).clone(), T(g["sd.bn.running_var"]).clone())
    return p, st


@pytest.mark.parametrize("tag", ["train_d32", "eval_d32", "train_d256", "mult_d32"])
def test_layer_matches_reference(golden, tag):
    g = golden("layer_" + tag)
    p, st = layer_from_golden(g)
    x = T(g["x"]).clone().requires_grad_(True)
    r = T(g["r"]).clone().requires_grad_(True)
    out = O.relation_aware_layer(p, x, r, T(g["edge_index"]), T(g["edge_type"]), slope=0.05,
                                 comp_op="mult" if "mult" in tag else "sub", bn_state=st,
                                 training=bool(g["training"]))
    np.testing.assert_allclose(out.detach().numpy(), g["out"], atol=1e-5, rtol=1e-4)
    (out * T(g["gout"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g["grad_x"], atol=2e-5, rtol=1e-3)
    np.testing.assert_allclose(r.grad.numpy(), g["grad_r"], atol=2e-5, rtol=1e-3)
    for k in O.LAYER_PARAM_NAMES:
        if k == "loop_rel" and bool(g["training"]):
            # mathematically zero in train mode (BN removes the constant shift); fp32 noise only
            np.testing.assert_allclose(p[k].grad.numpy(), g["grad." + k], atol=5e-4)
            continue
        scale = max(1.0, float(np.abs(g["grad." + k]).max()))
        np.testing.assert_allclose(p[k].grad.numpy(), g["grad." + k], atol=3e-5 * scale, rtol=1e-3, err_msg=k)
    if bool(g["training"]):
        np.testing.assert_allclose(st.running_mean.numpy(), g["after.bn.running_mean"], atol=1e-6, rtol=1e-5)
        np.testing.assert_allclose(st.running_var.numpy(), g["after.bn.running_var"], atol=1e-6, rtol=1e-5)


def model_sd(g):
    return {k[3:]: T(v).clone() for k, v in g.items() if k.startswith("sd.")}


@pytest.mark.parametrize("source", [True, False])
def test_completion_loss_matches_reference(golden, source):
    g = golden("model_noname")
    sd = model_sd(g)
    leaves = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    n1, n2, nr, B, K = (int(g[k]) for k in ("n1", "n2", "nr", "B", "K"))
    fd = {"links": g["links"], "ent_bases1": [0, n1], "ent_bases2": [n1, n1 + n2], "rel_bases1": [0, nr],
          "rel_bases2": [nr, 2 * nr]}
    h = np.tile(g["sub"], K + 1)
    r = np.tile(g["rel"], K + 1)
    t = np.concatenate([g["obj"], g["neg"].reshape(-1)])
    if not source:
        h, t = h % n2, t % n2
    data = {"batch_h": T(h), "batch_r": T(r), "batch_t": T(t)}
    loss = O.completion_loss(sd, data, T(g["ei1"]), T(g["et1"]), T(g["ei2"]), T(g["et2"]), fd, source,
                             batch_size=B, margin=5.0)
    tag = "src" if source else "dst"
    np.testing.assert_allclose(float(loss), float(g["closs_" + tag]), rtol=1e-5)
    loss.backward()
    for k, v in g.items():
        if k.startswith(f"cgrad_{tag}."):
            name = k.split(".", 1)[1]
            got = leaves[name].grad
            got = np.zeros_like(v) if got is None else got.numpy()
            np.testing.assert_allclose(got, v, atol=3e-5, rtol=1e-3, err_msg=name)


def test_alignment_loss_and_linkpred_match_reference(golden):
    g = golden("model_noname")
    sd = model_sd(g)
    n1, n2, nr, K = (int(g[k]) for k in ("n1", "n2", "nr", "K"))
    _, e1, _ = O.forward_no_name(sd, T(g["ei1"]), T(g["et1"]), [0, n1], [0, nr])
    _, e2, _ = O.forward_no_name(sd, T(g["ei2"]), T(g["et2"]), [n1, n1 + n2], [nr, 2 * nr])
    fa = {"links": g["links"], **{k: g["fa." + k] for k in ("neg_left", "neg_right", "neg2_left", "neg2_right")}}
    loss = O.alignment_margin_loss(e1[-1], e2[-1], fa, 1.0, K)
    np.testing.assert_allclose(float(loss), float(g["aloss"]), rtol=1e-5)
    # link prediction in eval mode
    sd = model_sd(g)
    _, ents, rels = O.forward_no_name(sd, T(g["ei1"]), T(g["et1"]), [0, n1], [0, nr], training=False)
    d = O.forward_linkpred(ents, rels, g["lp_e"], g["lp_r"])
    np.testing.assert_allclose(d.numpy(), g["lp_dist"], atol=1e-4, rtol=1e-5)
    dh = O.forward_linkpred(ents, rels, g["lp_e"], g["lp_r"], pred_head=True)
    np.testing.assert_allclose(dh.numpy(), g["lp_dist_head"], atol=1e-4, rtol=1e-5)
    emb = torch.nn.functional.normalize(ents[-1], 2, -1).numpy()
    np.testing.assert_allclose(emb, g["emb_comp"], atol=1e-6, rtol=1e-5)


def test_evaluator_matches_reference(golden):
    g = golden("evaluator")
    sd = model_sd(g)
    n, nr = int(g["n"]), int(g["nr"])
    _, ents, rels = O.forward_no_name(sd, T(g["ei"]), T(g["et"]), [0, n], [0, nr], training=False)
    allt = np.concatenate([g["train"], g["val"], g["test"]])
    true_tail = {}
    for h, r, t in allt:
        true_tail.setdefault((int(h), int(r)), set()).add(int(t))
    true_tail = {k: np.array(sorted(v)) for k, v in true_tail.items()}
    for is_val, split in ((1, g["val"]), (0, g["test"])):
        pred = -O.forward_linkpred(ents, rels, split[:, 0], split[:, 1])
        for filt in (1, 0):
            ranks = O.linkpred_ranks(pred, split[:, 0], split[:, 1], split[:, 2], true_tail if filt else None)
            np.testing.assert_allclose(O.rank_metrics(ranks), g[f"metrics_val{is_val}_filt{filt}"], rtol=1e-6)


def test_alignment_scorer_matches_reference(golden):
    g = golden("alignment")
    e1, e2 = g["e1"], g["e2"]
    np.testing.assert_allclose(O.sim(e1, e2, "cosine", False, 0), g["sim_cos"], atol=1e-6)
    np.testing.assert_allclose(O.sim(e1, e2, "cosine", False, 10), g["sim_cos_csls10"], atol=2e-6)
    np.testing.assert_allclose(O.sim(e1, e2, "inner", False, 5), g["sim_inner_csls5"], atol=2e-6)
    _, hits, mr, mrr, _, _ = O.greedy_alignment(e1, e2, (1, 5, 10), "cosine", False, 10)
    np.testing.assert_allclose(list(hits) + [mr, mrr], g["test_cos_csls10"], rtol=1e-6)
    _, hits, mr, mrr, _, _ = O.greedy_alignment(e1, e2, (1, 5, 10), "inner", False, 0)
    np.testing.assert_allclose(list(hits) + [mr, mrr], g["test_inner_nocsls"], rtol=1e-6)
    neg = O.get_neg(g["ill"], T(e1), T(e2), 7)
    assert np.array_equal(neg.numpy(), g["get_neg"])
    ent, s1, s2 = O.compute_alignment_quality(T(e1), T(e2), g["caq_l1"], g["caq_l2"])
    np.testing.assert_allclose(float(ent), float(g["caq_entropy"]), rtol=1e-5)
    np.testing.assert_allclose(s1.max(1)[0].numpy(), g["caq_rowmax"], rtol=1e-5)
    assert np.array_equal(s1.max(1)[1].numpy(), g["caq_rowarg"])
    np.testing.assert_allclose(s2.max(1)[0].numpy(), g["caq_rowmax2"], rtol=1e-5)


def test_scatter_restatement_against_naive_loops():
    rng = np.random.default_rng(0)
    idx = T(rng.integers(0, 9, 50))
    src = T(rng.standard_normal((50, 3)))
    out = O.scatter_sum(src, idx, 11)
    sm = O.scatter_softmax(src, idx, 11)
    for gidx in range(11):
        m = (idx == gidx).numpy()
        np.testing.assert_allclose(out[gidx].numpy(), src.numpy()[m].sum(0) if m.any() else 0, atol=1e-12)
        if m.any():
            s = src.numpy()[m]
            e = np.exp(s - s.max(0))
            np.testing.assert_allclose(sm.numpy()[m], e / e.sum(0), atol=1e-12)


def test_oracle_on_real_dbp5l_graphs(golden):
    """The oracle on the el/en graphs of the in-repo DBP-5L data (real hubs, isolated entities) against the
    reference's trace (oracle/gen_golden_dbp5l.py): seeded init checksums, eval-mode embeddings, the first
    alignment / completion losses."""
    from types import SimpleNamespace
    from jmac_b200.model import JMAC                                   # parameter container only: same init order as the reference
    g = golden("dbp5l_el_en")
    K, B = 25, 1000
    args = SimpleNamespace(dim=256, device="cpu", dropout=0.0, leaky_relu_w=0.05, comp_op="sub", margin_align=1.0, num_negative=K,
                           margin_completion=5.0, num_gcn_layer=2, no_name_info=True, batch_size=B)
    ne, nr = int(g["num_entities"]), int(g["num_relations"])
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in JMAC(args, np.zeros((ne, 300), dtype=np.float32), nr, ne).state_dict().items()}
    for k, v in sd.items():
        np.testing.assert_allclose([float(v.double().sum()), float(v.double().abs().sum())], g["init." + k], rtol=1e-9, atol=1e-9, err_msg=k)
    L = lambda k: T(g[k].astype(np.int64))
    eb1, rb1, eb2, rb2 = (g[k].tolist() for k in ("eb1", "rb1", "eb2", "rb2"))
    with torch.no_grad():
        _, ents, _ = O.forward_no_name(sd, L("ei1"), L("et1"), eb1, rb1, training=False)
        emb = torch.nn.functional.normalize(ents[-1], 2, -1).numpy()
        np.testing.assert_allclose(emb[::37], g["before.comp1_rows"], atol=1e-5, rtol=1e-4)
        # train mode (batch statistics), first alignment loss and first completion loss
        _, e1, _ = O.forward_no_name(sd, L("ei1"), L("et1"), eb1, rb1)
        _, e2, _ = O.forward_no_name(sd, L("ei2"), L("et2"), eb2, rb2)
        links = g["links"]
        fa = {"links": links, "neg_left": np.repeat(links[:, 0], K), "neg_right": g["neg_right"], "neg2_left": g["neg2_left"],
              "neg2_right": np.repeat(links[:, 1], K), "ent_bases1": eb1, "ent_bases2": eb2, "rel_bases1": rb1, "rel_bases2": rb2}
        loss = O.alignment_margin_loss(e1[-1], e2[-1], fa, 1.0, K)
        np.testing.assert_allclose(float(loss), g["align_loss"][0], rtol=1e-4)
