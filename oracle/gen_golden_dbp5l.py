"""Real-data golden trace (BASELINE.json configs[0]/[1]: in-repo DBP-5L, --no_name_info).

Runs the UNMODIFIED reference (imported from /root/reference through oracle/ref_shims.py) on the el/en pair
of datasetdbp5l: seeded model init, alignment evaluation, compute_alignment_quality, get_neg, three
optimisation rounds (alignment step + one completion batch per KG, Adam as in train.py:404-405), a final
alignment evaluation and the filtered link-prediction evaluation of the target KG.  The graphs, index
batches, negatives and every number the reference produced are written to tests/golden/dbp5l_el_en.npz;
tests/test_gpu_realdata.py replays the same procedure with jmac_b200 on a B200.

The model's parameters are NOT stored (58 MB): both sides initialise from torch.manual_seed(0) in the same
order (src/jmac_model.py:125-165), the fixture holds per-parameter checksums to prove it.

Build container only:   python oracle/gen_golden_dbp5l.py        TEST INFRASTRUCTURE ONLY.
"""
import logging
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "dbp5l_el_en.npz")
TARGET, OTHER = "el", "en"
ROUNDS, BATCH, K = 3, 1000, 25


def reference_args():
    """train.py:55-99 defaults, dropout off (CPU and CUDA dropout streams differ), CPU."""
    return SimpleNamespace(target_language=TARGET, data_path=os.path.join(ref_shims.REF, "datasetdbp5l"), margin_align=1.0,
                           margin_completion=5.0, dim=256, num_gcn_layer=2, align_percent=0.5, batch_size=BATCH, dropout=0.0,
                           leaky_relu_w=0.05, comp_op="sub", align_lr=3e-4, completion_lr=3e-4, num_negative=K,
                           pair_sample_weight=0.2, top_k=[1, 5, 10], csls=10, eval_metric="cosine", eval_norm=False,
                           noise_rate=0, device=torch.device("cpu"), no_name_info=True)


def checksums(model):
    return {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in model.state_dict().items()}


def main():
    logging.basicConfig(level=logging.WARNING)
    log = logging.getLogger("golden")
    ref_shims.install()
    ns = ref_shims.load_train_functions()
    from src.data_loader import ParseData
    from src.jmac_model import JMAC
    from src.validate import CompletionEvaluator
    from modules.utils.util import get_neg

    args = reference_args()
    rng = np.random.default_rng(0)
    ds = ParseData(args, log)
    kgs, seeds_train, seeds_test = ds.create_KG_objects_and_alignment()   # load_data minus the fastText name table (--no_name_info)
    num_entities, num_relations = ds.num_entities, kgs[TARGET].num_relation * ds.num_kgs
    pair = (TARGET, OTHER) if (TARGET, OTHER) in seeds_train else (OTHER, TARGET)
    kg1, kg2 = kgs[pair[0]], kgs[pair[1]]
    T = torch.LongTensor
    ei1, et1, ei2, et2 = T(kg1.edge_index), T(kg1.edge_type), T(kg2.edge_index), T(kg2.edge_type)
    eb1, rb1 = [kg1.entity_id_base, kg1.upper_entity_base], [kg1.relation_id_base, kg1.upper_relation_base]
    eb2, rb2 = [kg2.entity_id_base, kg2.upper_entity_base], [kg2.relation_id_base, kg2.upper_relation_base]
    links, test_links = seeds_train[pair], seeds_test[pair]
    res = {"pair": np.array(pair), "ei1": ei1.numpy().astype(np.int32), "et1": et1.numpy().astype(np.int32),
           "ei2": ei2.numpy().astype(np.int32), "et2": et2.numpy().astype(np.int32), "eb1": eb1, "rb1": rb1, "eb2": eb2, "rb2": rb2,
           "num_entities": num_entities, "num_relations": num_relations, "links": links, "test_links": test_links,
           "n1": kg1.num_entity, "n2": kg2.num_entity}

    torch.manual_seed(0)
    model = JMAC(args, np.zeros((num_entities, 300), dtype=np.float32), num_relations, num_entities)
    for k, v in checksums(model).items():
        res["init." + k] = v
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    align_opt = torch.optim.Adam(model.parameters(), lr=args.align_lr)
    comp_opt = torch.optim.Adam(model.parameters(), lr=args.completion_lr)

    def evaluate(tag):
        model.eval()
        with torch.no_grad():
            o1, c1 = model.get_emb(ei1, et1, eb1, rb1, pyt=True)
            o2, _ = model.get_emb(ei2, et2, eb2, rb2, pyt=True)
        _, hits, mr, mrr = ns["test_alignment_"](o1.numpy(), o2.numpy(), test_links[:, 0].tolist(), test_links[:, 1].tolist(), args, log)
        res[f"{tag}.align_metrics"] = np.array(list(hits) + [mr, mrr])
        res[f"{tag}.emb1_rows"] = o1[::37].numpy()
        res[f"{tag}.emb2_rows"] = o2[::101].numpy()
        res[f"{tag}.comp1_rows"] = c1[::37].numpy()
        return o1, o2

    # the seed-enlargement inputs of the first epoch (train.py:138-200)
    o1, o2 = evaluate("before")
    ent, s1, s2 = ns["compute_alignment_quality"](o1, o2, test_links[:, 0].tolist(), test_links[:, 1].tolist())
    res["caq_entropy"] = float(ent)
    res["caq_rowmax"], res["caq_rowarg"] = [t.numpy() for t in s1.max(1)]
    neg2_left = get_neg(links[:, 1].tolist(), o2, o1, K)
    neg_right = get_neg(links[:, 0].tolist(), o1, o2, K)
    res["neg2_left"], res["neg_right"] = np.asarray(neg2_left), np.asarray(neg_right)
    L = len(links)
    feed = {"neg_left": (np.ones((L, K)) * links[:, 0].reshape(L, 1)).reshape(-1), "neg_right": neg_right, "neg2_left": neg2_left,
            "neg2_right": (np.ones((L, K)) * links[:, 1].reshape(L, 1)).reshape(-1), "links": links, "ent_bases1": eb1, "ent_bases2": eb2,
            "rel_bases1": rb1, "rel_bases2": rb2}

    # completion batches with precomputed negatives (uniform over the KG's entities, shared list)
    tri1, tri2 = kg1.train_data, kg2.train_data
    batches = []
    for _ in range(ROUNDS):
        b = []
        for tri, n in ((tri1, kg1.num_entity), (tri2, kg2.num_entity)):
            sel = rng.choice(len(tri), BATCH, replace=False)
            b.append((tri[sel].astype(np.int64), rng.integers(0, n, (BATCH, K)).astype(np.int64)))
        batches.append(b)
    res["batches_tri"] = np.stack([np.stack([b[0][0], b[1][0]]) for b in batches])      # [ROUNDS, 2, B, 3]
    res["batches_neg"] = np.stack([np.stack([b[0][1], b[1][1]]) for b in batches])      # [ROUNDS, 2, B, K]

    def record_grads(tag):
        """the GNN layer's parameter gradients in full, every 97th row of the embedding tables', every 5th of the other matrices'"""
        for k, v in model.named_parameters():
            if v.grad is not None:
                step = 97 if v.shape[0] > 2048 else (1 if k.startswith("conv1_completion") or v.numel() <= 4096 else 5)
                res[f"{tag}.{k}"] = v.grad[::step].numpy().copy()

    model.train()
    aloss, closs = [], []
    for rnd in range(ROUNDS):
        align_opt.zero_grad()
        al = model.alignment_loss(feed, ei1, et1, ei2, et2)
        al.backward()
        if rnd == 0:
            record_grads("agrad")
        align_opt.step()
        aloss.append(float(al))
        for side, source in ((0, True), (1, False)):
            tri, neg = batches[rnd][side]
            sub, rel, obj = (T(tri[:, i]) for i in range(3))
            data = {"batch_h": sub.repeat(K + 1), "batch_r": rel.repeat(K + 1), "batch_t": torch.cat((obj, T(neg).view(-1)))}
            comp_opt.zero_grad()
            cl = model.completion_loss(data, ei1, et1, ei2, et2, feed, source)
            cl.backward()
            if rnd == 0 and side == 0:
                record_grads("cgrad")
            comp_opt.step()
            closs.append(float(cl))
        print("round", rnd, aloss[-1], closs[-2:], flush=True)
    res["align_loss"], res["completion_loss"] = np.array(aloss), np.array(closs)
    evaluate("after")

    # the same trace by the reference in float64 (same init, same inputs): |fp32 - fp64| of the reference itself
    # calibrates how far Adam (update = lr * g / (|g| + eps)) spreads fp32 rounding noise after nine steps
    m64 = JMAC(args, np.zeros((num_entities, 300), dtype=np.float32), num_relations, num_entities)
    m64.load_state_dict(sd0)
    m64 = m64.double()
    opt_a, opt_c = torch.optim.Adam(m64.parameters(), lr=args.align_lr), torch.optim.Adam(m64.parameters(), lr=args.completion_lr)
    m64.train()
    a64, c64 = [], []
    for rnd in range(ROUNDS):
        opt_a.zero_grad()
        al = m64.alignment_loss(feed, ei1, et1, ei2, et2)
        al.backward()
        opt_a.step()
        a64.append(float(al))
        for side, source in ((0, True), (1, False)):
            tri, neg = batches[rnd][side]
            sub, rel, obj = (T(tri[:, i]) for i in range(3))
            data = {"batch_h": sub.repeat(K + 1), "batch_r": rel.repeat(K + 1), "batch_t": torch.cat((obj, T(neg).view(-1)))}
            opt_c.zero_grad()
            cl = m64.completion_loss(data, ei1, et1, ei2, et2, feed, source)
            cl.backward()
            opt_c.step()
            c64.append(float(cl))
    m64.eval()
    with torch.no_grad():
        o1, c1 = m64.get_emb(ei1, et1, eb1, rb1, pyt=True)
    res["f64.align_loss"], res["f64.completion_loss"] = np.array(a64), np.array(c64)
    res["f64.after.emb1_rows"], res["f64.after.comp1_rows"] = o1[::37].numpy(), c1[::37].numpy()
    print("reference fp32 vs fp64 after training: emb", float(np.abs(res["after.emb1_rows"] - res["f64.after.emb1_rows"]).max()),
          "losses", np.abs(res["align_loss"] - a64).max(), np.abs(res["completion_loss"] - c64).max())

    # filtered link prediction on the target KG (src/validate.py:22-80)
    model.eval()
    tkg = kgs[TARGET]
    ev = CompletionEvaluator(tkg, model, "cpu", args.data_path)
    with torch.no_grad():
        for is_val in (True, False):
            res[f"lp_val{int(is_val)}"] = np.array(ev.test(args, is_val=is_val, filterr=True, logger=log))
    res["tkg_train"], res["tkg_val"], res["tkg_test"] = tkg.train_data, tkg.val_data, tkg.test_data
    res["tkg_edge_index"], res["tkg_edge_type"] = np.asarray(tkg.edge_index).astype(np.int32), np.asarray(tkg.edge_type).astype(np.int32)
    res["tkg_bases"] = np.array([tkg.entity_id_base, tkg.upper_entity_base, tkg.relation_id_base, tkg.upper_relation_base])
    np.savez_compressed(OUT, **{k: np.asarray(v) for k, v in res.items()})
    print({k: res[k] for k in ("before.align_metrics", "after.align_metrics", "align_loss", "completion_loss", "lp_val1", "lp_val0", "caq_entropy")})
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
