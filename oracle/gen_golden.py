"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference
through oracle/ref_shims.py) on small seeded inputs.  Run in the build container only:

    python oracle/gen_golden.py

The fixtures pin the oracle (tests/test_oracle_golden.py) and, on the GPU box, the CUDA path
(tests/test_gpu_*.py).  TEST INFRASTRUCTURE ONLY.
"""
import argparse
import logging
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_args(dim, **kw):
    a = SimpleNamespace(dim=dim, device="cpu", dropout=0.0, leaky_relu_w=0.05, comp_op="sub", margin_align=1.0,
                        num_negative=3, margin_completion=5.0, num_gcn_layer=2, no_name_info=True, batch_size=8)
    a.__dict__.update(kw)
    return a


def random_graph(rng, n, e, nr, hub=None, isolated=()):
    row = rng.integers(0, n, size=e)
    col = rng.integers(0, n, size=e)
    if hub is not None:
        row[: e // 3] = hub                       # a hub destination
    for z in isolated:                            # nodes with in-degree 0
        row[row == z] = (z + 1) % n
    # duplicate edges on purpose
    row[-3:] = row[:3]
    col[-3:] = col[:3]
    typ = rng.integers(0, nr, size=e)
    typ[-3:] = typ[:3]
    return np.stack([row, col]).astype(np.int64), typ.astype(np.int64)


def npify(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def gen_layer(tag, n, e, nr, d, seed, training=True, comp_op="sub"):
    from src.jmac_model import RelationAwareLayer
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    args = make_args(d, comp_op=comp_op)
    layer = RelationAwareLayer(d, d, rel_dim=d, act=torch.tanh, args=args)
    with torch.no_grad():                          # non-trivial BN affine + running stats
        layer.bn.weight.copy_(torch.rand(d) + 0.5)
        layer.bn.bias.copy_(torch.randn(d) * 0.1)
        layer.bn.running_mean.copy_(torch.randn(d) * 0.05)
        layer.bn.running_var.copy_(torch.rand(d) * 0.5 + 0.5)
    layer.train(training)
    ei, et = random_graph(rng, n, e, nr, hub=3, isolated=(0, n - 1))
    x = torch.randn(n, d, requires_grad=True)
    r = torch.randn(nr, d, requires_grad=True)
    gout = torch.randn(n, d)
    sd0 = {k: v.clone() for k, v in layer.state_dict().items()}
    out = layer(x, r, torch.from_numpy(ei), torch.from_numpy(et))
    (out * gout).sum().backward()
    res = {"edge_index": ei, "edge_type": et, "x": x, "r": r, "gout": gout, "out": out,
           "grad_x": x.grad, "grad_r": r.grad, "training": np.array(training)}
    for k, v in sd0.items():
        res["sd." + k] = v
    for k, v in layer.named_parameters():
        res["grad." + k] = v.grad if v.grad is not None else torch.zeros_like(v)
    for k, v in layer.state_dict().items():
        if "running" in k or "num_batches" in k:
            res["after." + k] = v
    np.savez_compressed(os.path.join(OUT, f"layer_{tag}.npz"), **npify(res))
    print("layer", tag, "out", float(out.abs().mean()))


def gen_model(seed=11):
    from src.jmac_model import JMAC
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    d, B, K = 16, 8, 3
    n1, n2, nr = 30, 22, 5                       # two KGs, global tables of 52 entities / 10 relations
    args = make_args(d, num_negative=K, batch_size=B)
    model = JMAC(args, np.zeros((n1 + n2, 300), dtype=np.float32), 2 * nr, n1 + n2)
    model.train()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    ei1, et1 = random_graph(rng, n1, 90, nr, hub=2, isolated=(0,))
    ei2, et2 = random_graph(rng, n2, 70, nr, hub=5, isolated=(1,))
    links = np.stack([rng.permutation(n1)[:6], rng.permutation(n2)[:6]], 1).astype(np.int64)
    fd = {"links": links, "ent_bases1": [0, n1], "ent_bases2": [n1, n1 + n2], "rel_bases1": [0, nr],
          "rel_bases2": [nr, 2 * nr]}
    # the batch layout of train.py:347-352
    sub, rel, obj = rng.integers(0, n1, B), rng.integers(0, nr, B), rng.integers(0, n1, B)
    neg = rng.integers(0, n1, (B, K))
    data = {"batch_h": torch.from_numpy(np.tile(sub, K + 1)), "batch_r": torch.from_numpy(np.tile(rel, K + 1)),
            "batch_t": torch.from_numpy(np.concatenate([obj, neg.reshape(-1)]))}
    t = lambda a: torch.from_numpy(a)
    res = {"ei1": ei1, "et1": et1, "ei2": ei2, "et2": et2, "links": links, "sub": sub, "rel": rel, "obj": obj,
           "neg": neg, "n1": n1, "n2": n2, "nr": nr, "B": B, "K": K}
    for k, v in sd0.items():
        res["sd." + k] = v
    # completion loss (source=True and False) + grads
    for src_flag in (True, False):
        model.load_state_dict(sd0)
        model.zero_grad()
        if not src_flag:                            # tails/heads must index KG2 (22 entities)
            d2 = {k: v % n2 for k, v in data.items() if k != "batch_r"}
            d2["batch_r"] = data["batch_r"]
        else:
            d2 = data
        loss = model.completion_loss(d2, t(ei1), t(et1), t(ei2), t(et2), fd, src_flag)
        loss.backward()
        tag = "src" if src_flag else "dst"
        res[f"closs_{tag}"] = loss
        for k, v in model.named_parameters():
            if v.grad is not None:
                res[f"cgrad_{tag}." + k] = v.grad.clone()
    # alignment loss + grads
    model.load_state_dict(sd0)
    model.zero_grad()
    L = len(links)
    fa = dict(fd)
    fa["neg_left"] = (np.ones((L, K)) * links[:, 0].reshape(L, 1)).reshape(-1)        # float64, train.py:189-190
    fa["neg_right"] = t(rng.integers(0, n2, L * K))
    fa["neg2_left"] = t(rng.integers(0, n1, L * K))
    fa["neg2_right"] = (np.ones((L, K)) * links[:, 1].reshape(L, 1)).reshape(-1)
    aloss = model.alignment_loss(fa, t(ei1), t(et1), t(ei2), t(et2))
    aloss.backward()
    res["aloss"] = aloss
    for k in ("neg_left", "neg_right", "neg2_left", "neg2_right"):
        res["fa." + k] = fa[k]
    for k, v in model.named_parameters():
        if v.grad is not None:
            res["agrad." + k] = v.grad.clone()
    # link prediction distances + get_emb in eval mode
    model.load_state_dict(sd0)
    model.eval()
    with torch.no_grad():
        e_idx, r_idx = rng.integers(0, n1, 7).tolist(), rng.integers(0, nr, 7).tolist()
        dist = model.forward_linkpred(e_idx, r_idx, t(ei1), t(et1), list(range(n1)), [0, n1], [0, nr])
        dist_head = model.forward_linkpred(e_idx, r_idx, t(ei1), t(et1), list(range(n1)), [0, n1], [0, nr], pred_head=True)
        emb_a, emb_c = model.get_emb(t(ei1), t(et1), [0, n1], [0, nr])
    res.update({"lp_e": e_idx, "lp_r": r_idx, "lp_dist": dist, "lp_dist_head": dist_head, "emb_align": emb_a,
                "emb_comp": emb_c})
    np.savez_compressed(os.path.join(OUT, "model_noname.npz"), **npify(res))
    print("model closs", float(res["closs_src"]), float(res["closs_dst"]), "aloss", float(aloss))


def gen_model_name(seed=13):
    """The name-information encoder (JMAC.forward_name, src/jmac_model.py:172-204; dropout off): three
    RelationAwareLayers, the name projection and the concat-linear fusions, with both losses and all gradients."""
    from src.jmac_model import JMAC
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    d, B, K = 32, 8, 3
    n1, n2, nr = 40, 28, 6
    args = make_args(d, num_negative=K, batch_size=B, no_name_info=False)
    name_emb = rng.standard_normal((n1 + n2, 300)).astype(np.float32) * 0.3
    model = JMAC(args, name_emb, 2 * nr, n1 + n2)
    model.train()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    ei1, et1 = random_graph(rng, n1, 130, nr, hub=2, isolated=(0,))
    ei2, et2 = random_graph(rng, n2, 90, nr, hub=5, isolated=(1,))
    links = np.stack([rng.permutation(n1)[:7], rng.permutation(n2)[:7]], 1).astype(np.int64)
    fd = {"links": links, "ent_bases1": [0, n1], "ent_bases2": [n1, n1 + n2], "rel_bases1": [0, nr], "rel_bases2": [nr, 2 * nr]}
    t = lambda a: torch.from_numpy(a)
    sub, rel, obj = rng.integers(0, n1, B), rng.integers(0, nr, B), rng.integers(0, n1, B)
    neg = rng.integers(0, n1, (B, K))
    data = {"batch_h": t(np.tile(sub, K + 1)), "batch_r": t(np.tile(rel, K + 1)), "batch_t": t(np.concatenate([obj, neg.reshape(-1)]))}
    res = {"name_emb": name_emb, "ei1": ei1, "et1": et1, "ei2": ei2, "et2": et2, "links": links, "sub": sub, "rel": rel, "obj": obj,
           "neg": neg, "n1": n1, "n2": n2, "nr": nr, "B": B, "K": K}
    for k, v in sd0.items():
        res["sd." + k] = v
    out, cembs, crels = model.forward_name(t(ei1), t(et1), [0, n1], [0, nr])
    res.update({"fwd_out": out, "fwd_comp1": cembs[1], "fwd_rel1": crels[1]})
    model.zero_grad()
    closs = model.completion_loss(data, t(ei1), t(et1), t(ei2), t(et2), fd, True)
    closs.backward()
    res["closs"] = closs
    for k, v in model.named_parameters():
        if v.grad is not None:
            res["cgrad." + k] = v.grad.clone()
    model.zero_grad()
    L = len(links)
    fa = dict(fd)
    fa["neg_left"] = (np.ones((L, K)) * links[:, 0].reshape(L, 1)).reshape(-1)
    fa["neg_right"] = t(rng.integers(0, n2, L * K))
    fa["neg2_left"] = t(rng.integers(0, n1, L * K))
    fa["neg2_right"] = (np.ones((L, K)) * links[:, 1].reshape(L, 1)).reshape(-1)
    aloss = model.alignment_loss(fa, t(ei1), t(et1), t(ei2), t(et2))
    aloss.backward()
    res["aloss"] = aloss
    for k in ("neg_left", "neg_right", "neg2_left", "neg2_right"):
        res["fa." + k] = fa[k]
    for k, v in model.named_parameters():
        if v.grad is not None:
            res["agrad." + k] = v.grad.clone()
    model.load_state_dict(sd0)
    model.eval()
    with torch.no_grad():
        emb_a, emb_c = model.get_emb(t(ei1), t(et1), [0, n1], [0, nr])
    res.update({"emb_align": emb_a, "emb_comp": emb_c})
    np.savez_compressed(os.path.join(OUT, "model_name.npz"), **npify(res))
    print("model(name) closs", float(closs), "aloss", float(aloss), "agrads", sum(k.startswith("agrad.") for k in res))


def gen_evaluator(seed=5):
    """CompletionEvaluator.test on a small KnowledgeGraph (src/validate.py:22-80)."""
    from src.jmac_model import JMAC
    from src.knowledgegraph import KnowledgeGraph
    from src.validate import CompletionEvaluator
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    d, n, nr = 16, 40, 6
    args = make_args(d, batch_size=16)
    tri = np.stack([rng.integers(0, n, 150), rng.integers(0, nr, 150), rng.integers(0, n, 150)], 1).astype(np.int64)
    train, val, test = tri[:90], tri[90:125], tri[125:]
    kg = KnowledgeGraph("xx", train, val, test, n, nr, False, 0, 0, "cpu")
    kg.upper_entity_base, kg.upper_relation_base = n, nr
    ei = np.stack([np.concatenate([train[:, 0], train[:, 2]]), np.concatenate([train[:, 2], train[:, 0]])])
    et = np.concatenate([train[:, 1], train[:, 1]])          # bidirectional, same rel id (src/utils.py:127-145)
    kg.edge_index, kg.edge_type = ei, et
    model = JMAC(args, np.zeros((n, 300), dtype=np.float32), nr, n)
    model.eval()
    ev = CompletionEvaluator(kg, model, "cpu", None)
    log = logging.getLogger("golden")
    res = {"train": train, "val": val, "test": test, "ei": ei, "et": et, "n": n, "nr": nr}
    for k, v in model.state_dict().items():
        res["sd." + k] = v
    with torch.no_grad():
        for is_val in (True, False):
            for filt in (True, False):
                h1, h10, mrr = ev.test(args, is_val=is_val, filterr=filt, logger=log)
                res[f"metrics_val{int(is_val)}_filt{int(filt)}"] = np.array([h1, h10, mrr])
    np.savez_compressed(os.path.join(OUT, "evaluator.npz"), **npify(res))
    print("evaluator", {k: v for k, v in res.items() if k.startswith("metrics")})


def gen_alignment(seed=3):
    from modules.finding.similarity import sim, csls_sim
    from modules.finding.evaluation import test as ref_test
    from modules.utils.util import get_neg
    rng = np.random.default_rng(seed)
    n, d = 96, 32
    e1 = rng.standard_normal((n, d)).astype(np.float32)
    e2 = (e1 + 2.0 * rng.standard_normal((n, d))).astype(np.float32)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 /= np.linalg.norm(e2, axis=1, keepdims=True)
    log = logging.getLogger("golden")
    res = {"e1": e1, "e2": e2}
    res["sim_cos"] = sim(e1, e2, metric="cosine", normalize=False, csls_k=0)
    res["sim_cos_csls10"] = sim(e1, e2, metric="cosine", normalize=False, csls_k=10)
    res["sim_inner_csls5"] = sim(e1, e2, metric="inner", normalize=False, csls_k=5)
    for name, kw in (("cos_csls10", dict(metric="cosine", csls_k=10)), ("inner_nocsls", dict(metric="inner", csls_k=0))):
        top_k, hits, mr, mrr = ref_test(e1, e2, None, [1, 5, 10], 1, normalize=False, accurate=True, logger=log, **kw)
        res["test_" + name] = np.array(list(hits) + [mr, mrr])
    ill = rng.permutation(n)[:20]
    res["ill"] = ill
    res["get_neg"] = get_neg(ill.tolist(), torch.from_numpy(e1), torch.from_numpy(e2), 7)
    ns = ref_shims.load_train_functions()
    l1, l2 = rng.permutation(n)[:30].tolist(), rng.permutation(n)[:30].tolist()
    ent, s1, s2 = ns["compute_alignment_quality"](torch.from_numpy(e1), torch.from_numpy(e2), l1, l2)
    res.update({"caq_l1": l1, "caq_l2": l2, "caq_entropy": ent, "caq_rowmax": s1.max(1)[0], "caq_rowarg": s1.max(1)[1],
                "caq_rowmax2": s2.max(1)[0], "caq_rowarg2": s2.max(1)[1]})
    np.savez_compressed(os.path.join(OUT, "alignment.npz"), **npify(res))
    print("alignment", res["test_cos_csls10"], res["test_inner_nocsls"])


if __name__ == "__main__":
    argparse.ArgumentParser().parse_args()
    logging.basicConfig(level=logging.WARNING)
    ref_shims.install()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    gen_layer("train_d32", n=60, e=300, nr=7, d=32, seed=1, training=True)
    gen_layer("eval_d32", n=60, e=300, nr=7, d=32, seed=2, training=False)
    gen_layer("train_d256", n=96, e=700, nr=11, d=256, seed=3, training=True)
    gen_layer("mult_d32", n=40, e=160, nr=5, d=32, seed=4, training=True, comp_op="mult")
    gen_model()
    gen_model_name()
    gen_evaluator()
    gen_alignment()
