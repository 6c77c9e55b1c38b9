"""Real-data parity (BASELINE.json configs[0]/[1]): replay of the reference's trace on the el/en pair of the in-repo
DBP-5L data (--no_name_info, dim 256) recorded by oracle/gen_golden_dbp5l.py from the unmodified reference:
seeded init, alignment evaluation, compute_alignment_quality, get_neg, three optimisation rounds (alignment
step + one completion batch per KG, Adam), final alignment evaluation, filtered link prediction on the target KG.
Real degree distribution: in-degree max 481-2,769 against a mean of 3-7, 6-37 % isolated entities.

Tolerances: fp32 1e-4 relative on losses/embeddings before any optimiser step; Adam amplifies rounding noise on
near-zero gradients (update = lr * g / (|g| + eps)), so quantities after k steps get k * lr-sized slack.
Alignment metrics, get_neg and compute_alignment_quality run in the default fp32-accurate (3xTF32) mode of the
similarity kernel: what separates them from the reference's numbers is only the ~1e-5 fp32 noise of the embeddings
themselves (documented ties)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
K, B = 25, 1000


def T(a):
    return torch.from_numpy(np.asarray(a).astype(np.int64))


def make_args():
    return SimpleNamespace(dim=256, device="cuda", dropout=0.0, leaky_relu_w=0.05, comp_op="sub", margin_align=1.0, num_negative=K,
                           margin_completion=5.0, num_gcn_layer=2, no_name_info=True, batch_size=B, top_k=[1, 5, 10], csls=10,
                           eval_metric="cosine", eval_norm=False, align_lr=3e-4, completion_lr=3e-4)


@pytest.fixture(scope="module")
def trace(golden):
    """Run the whole procedure once; the tests below look at different parts of it."""
    from jmac_b200 import JMAC, CompletionEvaluator
    from jmac_b200 import finding as F
    g = golden("dbp5l_el_en")
    args = make_args()
    ne, nr = int(g["num_entities"]), int(g["num_relations"])
    torch.manual_seed(0)
    model = JMAC(args, np.zeros((ne, 300), dtype=np.float32), nr, ne)
    init = {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in model.state_dict().items()}
    model = model.cuda()
    out = {"g": g, "init": init}
    ei1, et1, ei2, et2 = (T(g[k]).cuda() for k in ("ei1", "et1", "ei2", "et2"))
    eb1, rb1, eb2, rb2 = (g[k].tolist() for k in ("eb1", "rb1", "eb2", "rb2"))
    links, test_links = g["links"], g["test_links"]
    align_opt = torch.optim.Adam(model.parameters(), lr=args.align_lr)
    comp_opt = torch.optim.Adam(model.parameters(), lr=args.completion_lr)

    def evaluate(tag):
        model.eval()
        with torch.no_grad():
            o1, c1 = model.get_emb(ei1, et1, eb1, rb1, pyt=True)
            o2, _ = model.get_emb(ei2, et2, eb2, rb2, pyt=True)
        _, hits, mr, mrr = F.test_alignment_(o1, o2, test_links[:, 0].tolist(), test_links[:, 1].tolist(), args, None)
        out[tag] = {"metrics": np.array(list(hits) + [mr, mrr]), "emb1": o1, "emb2": o2, "comp1": c1}
        return o1, o2

    o1, o2 = evaluate("before")
    ent, s1, _ = F.compute_alignment_quality(o1, o2, test_links[:, 0].tolist(), test_links[:, 1].tolist())
    out["caq"] = (float(ent), *[t.cpu().numpy() for t in s1.max(dim=1)])
    out["neg_right"] = F.get_neg(links[:, 0].tolist(), o1, o2, K).cpu().numpy()
    L = len(links)
    # the reference's own negatives, so that the optimisation trace sees identical inputs
    feed = {"neg_left": (np.ones((L, K)) * links[:, 0].reshape(L, 1)).reshape(-1), "neg_right": T(g["neg_right"]),
            "neg2_left": T(g["neg2_left"]), "neg2_right": (np.ones((L, K)) * links[:, 1].reshape(L, 1)).reshape(-1), "links": links,
            "ent_bases1": eb1, "ent_bases2": eb2, "rel_bases1": rb1, "rel_bases2": rb2}
    model.train()
    aloss, closs, grads = [], [], {}

    def record(tag):
        for k, v in model.named_parameters():
            if v.grad is not None:
                grads[f"{tag}.{k}"] = v.grad.detach().cpu().numpy().copy()

    for rnd in range(len(g["align_loss"])):
        align_opt.zero_grad()
        al = model.alignment_loss(feed, ei1, et1, ei2, et2)
        al.backward()
        if rnd == 0:
            record("agrad")
        align_opt.step()
        aloss.append(float(al))
        for side, source in ((0, True), (1, False)):
            tri, neg = g["batches_tri"][rnd, side], g["batches_neg"][rnd, side]
            sub, rel, obj = (T(tri[:, i]).cuda() for i in range(3))
            data = {"batch_h": sub.repeat(K + 1), "batch_r": rel.repeat(K + 1), "batch_t": torch.cat((obj, T(neg).cuda().view(-1)))}
            comp_opt.zero_grad()
            cl = model.completion_loss(data, ei1, et1, ei2, et2, feed, source)
            cl.backward()
            if rnd == 0 and side == 0:
                record("cgrad")
            comp_opt.step()
            closs.append(float(cl))
    out["align_loss"], out["completion_loss"], out["grads"] = np.array(aloss), np.array(closs), grads
    evaluate("after")

    # filtered link prediction on the target KG through the drop-in evaluator
    allt = np.concatenate([g["tkg_train"], g["tkg_val"], g["tkg_test"]])
    true_tail = {}
    for h, r, t in allt:
        true_tail.setdefault((int(h), int(r)), set()).add(int(t))

    def gen(h, r, t, batch_size, shuffle=False):
        tri = torch.stack([h, r, t], 1)
        return [tri[i:i + batch_size] for i in range(0, len(tri), batch_size)]

    b = g["tkg_bases"].tolist()
    kg = SimpleNamespace(edge_index=g["tkg_edge_index"].astype(np.int64), edge_type=g["tkg_edge_type"].astype(np.int64), entity_id_base=b[0],
                         upper_entity_base=b[1], relation_id_base=b[2], upper_relation_base=b[3], num_entity=b[1] - b[0],
                         true_tail={k: np.array(sorted(v)) for k, v in true_tail.items()},
                         h_val=T(g["tkg_val"][:, 0]), r_val=T(g["tkg_val"][:, 1]), t_val=T(g["tkg_val"][:, 2]),
                         h_test=T(g["tkg_test"][:, 0]), r_test=T(g["tkg_test"][:, 1]), t_test=T(g["tkg_test"][:, 2]), generate_batch_data=gen)
    model.eval()
    ev = CompletionEvaluator(kg, model, "cuda", None)
    with torch.no_grad():
        out["lp"] = {v: np.array(ev.test(args, is_val=bool(v), filterr=True)) for v in (1, 0)}
    return out


def test_seeded_init_is_the_reference_init(trace):
    g = trace["g"]
    for k, v in trace["init"].items():
        np.testing.assert_allclose(v, g["init." + k], rtol=1e-9, atol=1e-9, err_msg=k)


def test_embeddings_before_training(trace):
    g, t = trace["g"], trace["before"]
    np.testing.assert_allclose(t["emb1"][::37].cpu().numpy(), g["before.emb1_rows"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(t["emb2"][::101].cpu().numpy(), g["before.emb2_rows"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(t["comp1"][::37].cpu().numpy(), g["before.comp1_rows"], atol=1e-5, rtol=1e-4)


def test_first_step_losses_and_gradients(trace):
    g = trace["g"]
    np.testing.assert_allclose(trace["align_loss"][0], g["align_loss"][0], rtol=1e-4)
    np.testing.assert_allclose(trace["completion_loss"][0], g["completion_loss"][0], rtol=1e-4)
    checked = 0
    for key, ref in g.items():
        if not (key.startswith("agrad.") or key.startswith("cgrad.")):
            continue
        got = trace["grads"].get(key)
        name = key.split(".", 1)[1]
        if got is None:
            got = np.zeros_like(ref)
        elif got.shape != ref.shape:                                   # the fixture holds every 97th / 5th row of the big tables
            step = 97 if got.shape[0] > 2048 else 5
            got = got[::step]
        scale = max(float(np.abs(ref).max()), 1e-6)
        if name.endswith("loop_rel"):                                    # mathematically zero in train mode (BatchNorm removes constants)
            assert float(np.abs(got).max()) <= 1e-4, key
        else:
            assert float(np.abs(got - ref).max()) <= 2e-3 * scale + 1e-7, f"{key}: {float(np.abs(got - ref).max()):.3e} vs scale {scale:.3e}"
        checked += 1
    assert checked >= 20


def test_loss_trace_over_adam_rounds(trace):
    g = trace["g"]
    np.testing.assert_allclose(trace["align_loss"], g["align_loss"], rtol=5e-3)
    np.testing.assert_allclose(trace["completion_loss"], g["completion_loss"], rtol=5e-3)
    assert trace["align_loss"][-1] < trace["align_loss"][0]              # it trains
    # after nine Adam steps: budget calibrated on the reference itself, run in fp32 and in fp64 from the same init
    # (its own fp32-vs-fp64 spread is 1.1e-2 on these unit-norm rows: Adam turns rounding noise on near-zero
    # gradients into lr-sized steps); the new path must sit as close to the fp64 trace as the reference's fp32 does
    t = trace["after"]
    for ours, key in ((t["emb1"], "after.emb1_rows"), (t["comp1"], "after.comp1_rows")):
        ref32, ref64 = g[key], g["f64." + key]
        err_ref = float(np.abs(ref32 - ref64).max())
        err_new = float(np.abs(ours[::37].cpu().numpy() - ref64).max())
        assert err_new <= 2.0 * err_ref + 1e-4, f"{key}: {err_new:.3e} vs reference fp32 {err_ref:.3e}"
        rms_ref = float(np.sqrt(np.mean((ref32 - ref64) ** 2)))
        rms_new = float(np.sqrt(np.mean((ours[::37].cpu().numpy() - ref64) ** 2)))
        assert rms_new <= 2.0 * rms_ref + 1e-5, f"{key}: rms {rms_new:.3e} vs reference fp32 {rms_ref:.3e}"
    np.testing.assert_allclose(trace["align_loss"], g["f64.align_loss"], rtol=2e-3)
    np.testing.assert_allclose(trace["completion_loss"], g["f64.completion_loss"], rtol=2e-3)


def test_alignment_metrics_and_negatives(trace):
    g = trace["g"]
    n = len(g["test_links"])
    for tag in ("before", "after"):
        got, ref = trace[tag]["metrics"], g[f"{tag}.align_metrics"]
        # before training the inputs agree to ~1e-5 (fp32 encoder noise): at most 2 rows may cross a Hits@k boundary; after nine Adam
        # steps the reference's own run-to-run spread is 2 rows (threaded scatter order) and Adam spreads rounding noise (see above)
        rows = 2 if tag == "before" else 10
        assert np.all(np.abs(got[:3] - ref[:3]) <= 100.0 * rows / n + 1e-9), (tag, got, ref)    # rows crossing a Hits@k boundary
        mr_tol, mrr_tol = (2e-3, 5e-3) if tag == "before" else (0.02, 0.1)
        assert abs(got[3] - ref[3]) <= mr_tol * ref[3] and abs(got[4] - ref[4]) <= mrr_tol * ref[4] + 1e-4, (tag, got, ref)
    # get_neg: the reference's neighbours and ours are the same sets wherever the 25th and 26th neighbour are further apart than
    # the fp32 noise of the embeddings; the similarity values of the chosen neighbours agree to that noise
    o1 = trace["before"]["emb1"].double().cpu().numpy()
    o2 = trace["before"]["emb2"].double().cpu().numpy()
    src = g["links"][:, 0]
    S = o1[src] @ o2.T
    ours_i, refs_i = trace["neg_right"].reshape(-1, K), g["neg_right"].reshape(-1, K)
    ours = np.sort(S[np.arange(len(src))[:, None], ours_i], axis=1)
    refs = np.sort(S[np.arange(len(src))[:, None], refs_i], axis=1)
    scale = float(np.abs(S).max())
    band = 3e-5 * scale
    assert float(np.abs(ours - refs).max()) <= band
    # identical index SETS modulo documented ties: an index chosen by only one side must score within the band of the row's 25th value
    kth = -np.sort(-S, axis=1)[:, K - 1]
    rows = np.arange(len(src))[:, None]
    in_ours = np.zeros_like(S, dtype=bool)
    in_refs = np.zeros_like(S, dtype=bool)
    in_ours[rows, ours_i] = True
    in_refs[rows, refs_i] = True
    only_one = in_ours ^ in_refs
    assert np.all(np.abs(S - kth[:, None])[only_one] <= band)
    assert (only_one.sum(1) == 0).mean() > 0.9                                              # and most rows agree outright
    ent, rowmax, rowarg = trace["caq"]
    np.testing.assert_allclose(ent, float(g["caq_entropy"]), rtol=5e-4)
    np.testing.assert_allclose(rowmax, g["caq_rowmax"], rtol=2e-3, atol=1e-6)
    sure = g["caq_rowmax"] > 0.5                                                            # an argmax holding most of the mass cannot flip
    assert np.array_equal(rowarg[sure], g["caq_rowarg"][sure])


def test_filtered_link_prediction_metrics(trace):
    g = trace["g"]
    for v in (1, 0):
        got, ref = trace["lp"][v], g[f"lp_val{v}"]
        n = len(g["tkg_val"] if v else g["tkg_test"])
        assert abs(got[0] - ref[0]) <= 3.0 / n and abs(got[1] - ref[1]) <= 6.0 / n, (v, got, ref)     # rows inside the 9-step Adam slack
        assert abs(got[2] - ref[2]) <= 0.03 * ref[2] + 1e-4, (v, got, ref)
