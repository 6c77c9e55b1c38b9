"""CPU oracle for JMAC's hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (plain torch-on-CPU / numpy, no custom kernels) of the reference's
algorithm for the three parts of the hot path named in BASELINE.json: the relation-aware GNN layer,
the completion (TransE-L1 margin) scorer and the alignment scorer.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it; the product package ``jmac_b200`` never does.

It deliberately keeps the reference's *un-factored* per-edge formulation (materialised [E,d]
gathers, the [E,2d] concat GEMM, scatter-softmax, scatter-sum) so that it is an independent check of
the restructured CUDA path, and it runs in any float dtype (fp64 is used as the arbiter for
gradient tolerances).

Parity pinning: the reference ships no tests and no golden vectors.  The
oracle is pinned against outputs of the reference itself, imported unmodified in the build
container by ``oracle/gen_golden.py`` (synthetic cases) and ``oracle/gen_golden_dbp5l.py`` (a trace on the
el/en graphs of the in-repo DBP-5L data), fixtures in ``tests/golden/*.npz``; see
``tests/test_oracle_golden.py``.

Third-party arithmetic restated here (absent from /root/reference): ``torch_scatter``
(``scatter_add``, ``scatter(reduce='sum')``, ``scatter_softmax``; PyPI ``torch-scatter``, unpinned
in JMAC_DBPv1/requirements.txt) and scipy's ``cdist(metric='cosine')`` (float64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# torch_scatter semantics (call sites: src/jmac_model.py:105, modules/helper/message_passing.py:24,28)
# --------------------------------------------------------------------------------------------
def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """``torch_scatter.scatter(src, index, dim=0, dim_size=dim_size, reduce='sum')``: rows of
    ``src`` are summed into ``out[index[e]]``; groups with no member stay zero."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    if src.numel():
        out.index_add_(0, index, src)
    return out


def scatter_softmax(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """``torch_scatter.scatter_softmax(src, index, dim=0)``: softmax over the entries sharing an
    index, with the per-group maximum subtracted first (composite/softmax.py upstream)."""
    if src.numel() == 0:
        return src.clone()
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    gmax = torch.full((dim_size,) + tuple(src.shape[1:]), -math.inf, dtype=src.dtype, device=src.device)
    gmax = gmax.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    shifted = src - gmax[index]
    ex = shifted.exp()
    denom = scatter_sum(ex, index, dim_size)
    return ex / denom[index]


def leaky_relu(x: torch.Tensor, slope: float) -> torch.Tensor:
    return torch.where(x > 0, x, x * slope)


# --------------------------------------------------------------------------------------------
# RelationAwareLayer (src/jmac_model.py:10-109 + modules/helper/message_passing.py:4-29,55-90)
# --------------------------------------------------------------------------------------------
LAYER_PARAM_NAMES = ("rel_transform_weight1", "rel_transform_weight2", "gcn_weight", "loop_rel",
                     "w_att", "a_att", "bn.weight", "bn.bias")


@dataclass
class BNState:
    """nn.BatchNorm1d buffers (src/jmac_model.py:28)."""
    running_mean: torch.Tensor
    running_var: torch.Tensor
    num_batches_tracked: int = 0
    eps: float = 1e-5
    momentum: float = 0.1


def compute_norm(row: torch.Tensor, num_ent: int, dtype) -> torch.Tensor:
    """src/jmac_model.py:99-109 -- deg^-1/2 of the aggregation target of every edge (inf -> 0)."""
    deg = scatter_sum(torch.ones(row.shape[0], dtype=dtype, device=row.device), row, num_ent)
    deg_inv = deg.pow(-0.5)
    deg_inv[deg_inv == float("inf")] = 0
    return deg_inv[row]


def propagate(p: Dict[str, torch.Tensor], x, rel_embed, edge_index, edge_type, edge_norm, num_ent,
              slope: float, comp_op: str) -> torch.Tensor:
    """MessagePassing.propagate (message_passing.py:55-90) -> RelationAwareLayer.message
    (jmac_model.py:80-89) -> scatter_ (message_passing.py:4-29).  ``edge_index[0]`` is both the
    ``_i`` gather index and the aggregation index."""
    x_i = x[edge_index[0]]
    x_j = x[edge_index[1]]
    rel = rel_embed.index_select(0, edge_type)
    if comp_op == "sub":                                   # jmac_model.py:61-62
        msg = x_j - rel
    elif comp_op == "mult":                                # jmac_model.py:63-64
        msg = x_j * rel
    else:                                                  # jmac_model.py:65-66
        raise NotImplementedError
    combined = torch.cat((x_i, msg), dim=1)                # jmac_model.py:75
    score = leaky_relu(combined @ p["w_att"], slope) @ p["a_att"]      # :76  -> [E,1]
    out = msg @ p["gcn_weight"]                            # :88
    alpha = scatter_softmax(score, edge_index[0], num_ent)  # message_passing.py:24
    if edge_norm is not None:
        alpha = alpha / edge_norm.view(-1, 1)              # :25-26  (a DIVISION by deg^-1/2)
    return scatter_sum(out * alpha, edge_index[0], num_ent)  # :27-28


def batchnorm(x: torch.Tensor, weight, bias, state: Optional[BNState], training: bool) -> torch.Tensor:
    """nn.BatchNorm1d forward over the node dimension (jmac_model.py:52)."""
    if training:
        mean = x.mean(0)
        var_b = x.var(0, unbiased=False)
        if state is not None:
            n = x.shape[0]
            with torch.no_grad():
                unb = var_b * (n / max(n - 1, 1))
                state.running_mean.mul_(1 - state.momentum).add_(state.momentum * mean.detach().to(state.running_mean.dtype))
                state.running_var.mul_(1 - state.momentum).add_(state.momentum * unb.detach().to(state.running_var.dtype))
                state.num_batches_tracked += 1
    else:
        mean = state.running_mean.to(x.dtype)
        var_b = state.running_var.to(x.dtype)
    eps = state.eps if state is not None else 1e-5
    return (x - mean) / torch.sqrt(var_b + eps) * weight + bias


def relation_aware_layer(p: Dict[str, torch.Tensor], ent_emb, rel_emb, edge_index, edge_type, *,
                         slope: float = 0.05, comp_op: str = "sub", bn_state: Optional[BNState] = None,
                         training: bool = True, rel_act: str = "leaky") -> torch.Tensor:
    """RelationAwareLayer.forward (src/jmac_model.py:33-53).  ``rel_act='relu'`` is the
    JMAC_DBPv1 variant (JMAC_DBPv1/models/jmac_model.py:27,51)."""
    num_ent = ent_emb.shape[0]
    rel = torch.cat([rel_emb, p["loop_rel"]], dim=0)                       # :39
    rel = rel @ p["rel_transform_weight1"]                                  # :40
    rel = leaky_relu(rel, slope) if rel_act == "leaky" else torch.relu(rel)  # :41
    rel = rel @ p["rel_transform_weight2"]                                  # :42
    loop = torch.arange(num_ent, device=ent_emb.device)
    loop_index = torch.stack([loop, loop])                                  # :44
    loop_type = torch.full((num_ent,), rel.shape[0] - 1, dtype=torch.long, device=ent_emb.device)  # :45
    norm = compute_norm(edge_index[0], num_ent, ent_emb.dtype)              # :47
    neigh = propagate(p, ent_emb, rel, edge_index, edge_type, norm, num_ent, slope, comp_op)   # :49
    selfm = propagate(p, ent_emb, rel, loop_index, loop_type, None, num_ent, slope, comp_op)   # :50
    pre = (neigh + selfm) / 2
    return torch.tanh(batchnorm(pre, p["bn.weight"], p["bn.bias"], bn_state, training))        # :52


# --------------------------------------------------------------------------------------------
# JMAC encoders and losses (src/jmac_model.py:172-380)
# --------------------------------------------------------------------------------------------
def layer_params(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k: sd[f"{prefix}.{k}"] for k in LAYER_PARAM_NAMES}


def bn_state_of(sd: Dict[str, torch.Tensor], prefix: str) -> BNState:
    return BNState(sd[f"{prefix}.bn.running_mean"], sd[f"{prefix}.bn.running_var"],
                   int(sd.get(f"{prefix}.bn.num_batches_tracked", 0)))


def forward_no_name(sd, edge_index, edge_type, ent_bases, rel_bases, *, slope=0.05, comp_op="sub",
                    num_gcn_layer=2, training=True, bn_states=None):
    """JMAC.forward_no_name (src/jmac_model.py:207-220)."""
    e0 = sd["ent_init_att_completion"][ent_bases[0]:ent_bases[1]]
    r0 = sd["rel_init_att_completion"][rel_bases[0]:rel_bases[1]]
    ents, rels = [e0], [r0]
    if num_gcn_layer == 2:
        st = (bn_states or {}).get("conv1_completion") if bn_states is not None else None
        if st is None:
            st = bn_state_of(sd, "conv1_completion")
        e1 = relation_aware_layer(layer_params(sd, "conv1_completion"), e0, r0, edge_index, edge_type,
                                  slope=slope, comp_op=comp_op, bn_state=st, training=training)
        r1 = leaky_relu(r0 @ sd["rel_linear11"], slope) @ sd["rel_linear12"]       # :216
        ents.append(e1)
        rels.append(r1)
    return ents[-1], ents, rels


def alignment_loss_simple(links, emb1, emb2):
    """JMAC.alignment_loss_simple (src/jmac_model.py:237-249)."""
    if not len(links):
        return 0
    links = torch.as_tensor(np.asarray(links)).long().to(emb1.device)
    a = F.normalize(emb1[links[:, 0]], 2, -1)
    b = F.normalize(emb2[links[:, 1]], 2, -1)
    return (1 - (a * b).sum(1)).mean()


def transe_margin(ent, rel, batch_h, batch_r, batch_t, batch_size: int, margin: float):
    """Scorer part of JMAC.completion_loss for one layer (src/jmac_model.py:345-374): L1 TransE
    score over the (1+K)*B triples, positives = first B, negatives = the rest viewed [K,B]^T."""
    score = ((ent[batch_h] + rel[batch_r]) - ent[batch_t]).abs().sum(-1).flatten()
    pos = score[:batch_size]
    pos = pos.view(-1, min(batch_size, len(pos))).permute(1, 0)
    neg = score[batch_size:]
    neg = neg.view(-1, min(batch_size, len(neg))).permute(1, 0)
    m = torch.tensor([margin], dtype=score.dtype, device=score.device)
    return torch.max(pos - neg, -m).mean() + m, score


def completion_loss(sd, data, ei1, et1, ei2, et2, feeddict, source=True, *, batch_size, margin=5.0,
                    slope=0.05, comp_op="sub", num_gcn_layer=2, training=True):
    """JMAC.completion_loss (src/jmac_model.py:316-380), no-name encoder."""
    _, ents1, rels1 = forward_no_name(sd, ei1, et1, feeddict["ent_bases1"], feeddict["rel_bases1"],
                                      slope=slope, comp_op=comp_op, num_gcn_layer=num_gcn_layer, training=training)
    _, ents2, rels2 = forward_no_name(sd, ei2, et2, feeddict["ent_bases2"], feeddict["rel_bases2"],
                                      slope=slope, comp_op=comp_op, num_gcn_layer=num_gcn_layer, training=training)
    loss = 0
    for layer in range(num_gcn_layer):
        ent, rel = (ents1[layer], rels1[layer]) if source else (ents2[layer], rels2[layer])
        l_res, _ = transe_margin(ent, rel, data["batch_h"], data["batch_r"], data["batch_t"], batch_size, margin)
        loss = loss + l_res + alignment_loss_simple(feeddict["links"], ents1[layer], ents2[layer])
    return loss


def alignment_margin_loss(emb1, emb2, feeddict, margin_align: float, k: int):
    """Loss part of JMAC.alignment_loss given the two encoder outputs (src/jmac_model.py:265-292)."""
    links = torch.as_tensor(np.asarray(feeddict["links"])).long()
    n = len(links)

    def as_idx(x):
        return torch.as_tensor(np.asarray(x)).reshape(-1).long()

    a = F.normalize(emb1[links[:, 0]], 2, -1)
    b = F.normalize(emb2[links[:, 1]], 2, -1)
    A = 1 - (a * b).sum(1)
    D = A + margin_align
    total = 0
    for left, right in (("neg_left", "neg_right"), ("neg2_left", "neg2_right")):
        nl = F.normalize(emb1[as_idx(feeddict[left])], 2, -1)
        nr = F.normalize(emb2[as_idx(feeddict[right])], 2, -1)
        B = 1 - (nl * nr).sum(1)
        total = total + torch.relu(-B.view(n, -1) + D.view(n, 1)).sum()
    return total / (2 * k * n)


def forward_linkpred(ents: Sequence[torch.Tensor], rels: Sequence[torch.Tensor], e_index, r_index,
                     pred_head=False) -> torch.Tensor:
    """Distance part of JMAC.forward_linkpred (src/jmac_model.py:301-313) with
    ``all_index = range(N)``: sum over layers of L1 cdist(e +/- r, all entities)."""
    e_index = torch.as_tensor(e_index).long()
    r_index = torch.as_tensor(r_index).long()
    dist = 0
    for ent, rel in zip(ents, rels):
        er = ent[e_index] - rel[r_index] if pred_head else ent[e_index] + rel[r_index]
        dist = dist + torch.cdist(er, ent, p=1)                     # :312
    return dist


def linkpred_ranks(predictions: torch.Tensor, h, r, t, true_tail: Optional[dict]) -> np.ndarray:
    """CompletionEvaluator.test inner loop (src/validate.py:50-70).  ``predictions`` = -dist.
    Returns 1-based ranks; filtered setting when ``true_tail`` is given (sentinel -1e6, :56)."""
    pred = predictions.clone()
    t = [int(x) for x in t]
    if true_tail is not None:
        for j in range(pred.shape[0]):
            filt = true_tail[(int(h[j]), int(r[j]))]
            tv = pred[j, t[j]].item()
            pred[j, torch.as_tensor(np.asarray(filt)).long()] = -1e6
            pred[j, t[j]] = tv
    _, sort_idx = torch.sort(pred, dim=1, descending=True)
    sort_idx = sort_idx.numpy()
    return np.array([int(np.where(sort_idx[j] == t[j])[0][0]) + 1 for j in range(pred.shape[0])])


def rank_metrics(ranks: np.ndarray) -> Tuple[float, float, float]:
    """hits@1, hits@10, mrr as validate.py:72-80 computes them."""
    ranks = np.asarray(ranks)
    return float(np.mean(ranks <= 1)), float(np.mean(ranks <= 10)), float(np.mean(1.0 / ranks))


# --------------------------------------------------------------------------------------------
# Alignment scorer (modules/finding/similarity.py, alignment.py, utils/util.py, train.py:231-259)
# --------------------------------------------------------------------------------------------
def cosine_sim_f64(e1: np.ndarray, e2: np.ndarray) -> np.ndarray:
    """``1 - scipy.spatial.distance.cdist(e1, e2, 'cosine')`` then ``astype(float32)``
    (similarity.py:44-46): float64 dot / (|u| |v|)."""
    a = np.asarray(e1, dtype=np.float64)
    b = np.asarray(e2, dtype=np.float64)
    na = np.sqrt((a * a).sum(1))
    nb = np.sqrt((b * b).sum(1))
    return (1.0 - (1.0 - (a @ b.T) / (na[:, None] * nb[None, :]))).astype(np.float32)


def nearest_k_mean(sim_mat: np.ndarray, k: int) -> np.ndarray:
    """calculate_nearest_k (similarity.py:81-84).  Spec = mean of the true k largest per row
    (np.partition(-S, k+1)[:, :k] returned exactly those in the survey's probe)."""
    part = -np.partition(-sim_mat, k + 1, axis=1)
    return np.mean(part[:, 0:k], axis=1)


def csls_sim(sim_mat: np.ndarray, k: int) -> np.ndarray:
    """csls_sim (similarity.py:58-78): 2*S - r1[i] - r2[j]."""
    r1 = nearest_k_mean(sim_mat, k)
    r2 = nearest_k_mean(sim_mat.T, k)
    out = 2 * sim_mat.T - r1
    return out.T - r2


def sim(e1, e2, metric="inner", normalize=False, csls_k=0) -> np.ndarray:
    """sim (similarity.py:13-55); metrics used by the reference: 'inner' and 'cosine'."""
    e1 = np.asarray(e1)
    e2 = np.asarray(e2)
    if normalize:
        e1 = e1 / np.maximum(np.linalg.norm(e1, axis=1, keepdims=True), 1e-12)
        e2 = e2 / np.maximum(np.linalg.norm(e2, axis=1, keepdims=True), 1e-12)
    if metric == "inner" or (metric == "cosine" and normalize):
        s = np.matmul(e1, e2.T)
    elif metric == "cosine":
        s = cosine_sim_f64(e1, e2)
    else:
        raise NotImplementedError(metric)
    if csls_k > 0:
        s = csls_sim(s, csls_k)
    return s


def greedy_alignment(e1, e2, top_k=(1, 5, 10), metric="cosine", normalize=False, csls_k=10,
                     sim_mat: Optional[np.ndarray] = None):
    """greedy_alignment + calculate_rank with accurate=True (alignment.py:10-69,87-112): gold
    column of row i is i; rank = position of gold in the descending argsort of the row.
    Returns (top_k, hits% rounded to 3 dp, mr, mrr, ranks[0-based], top1 index per row)."""
    s = sim(e1, e2, metric, normalize, csls_k) if sim_mat is None else sim_mat
    n = s.shape[0]
    ranks = np.empty(n, dtype=np.int64)
    top1 = np.empty(n, dtype=np.int64)
    for i in range(n):
        order = (-s[i, :]).argsort()
        top1[i] = order[0]
        ranks[i] = np.where(order == i)[0][0]
    hits = np.array([np.sum(ranks < k) for k in top_k]) / n * 100
    hits = np.array([round(h, 3) for h in hits])
    mr = float(np.sum(ranks + 1) / n)
    mrr = float(np.sum(1.0 / (ranks + 1)) / n)
    return list(top_k), hits, mr, mrr, ranks, top1


def greedy_alignment_torch(e1: torch.Tensor, e2: torch.Tensor, top_k=(1, 5, 10), csls_k=10):
    """The same evaluation (metric='cosine', accurate=True) as dense torch ops on the tensors' device: float64 cosine
    (cdist semantics, similarity.py:44-46) -> float32, CSLS from the means of the k largest per row / column
    (similarity.py:58-84), rank of the gold column in a full descending argsort of every row (alignment.py:99-104).
    This is the "reference algorithm in PyTorch eager" that bench.py times on the GPU next to the streaming kernel;
    it materialises the n1 x n2 matrix like the reference does.  Returns (hits%, mr, mrr, ranks[0-based])."""
    a, b = e1.double(), e2.double()
    s = (1.0 - (1.0 - (a @ b.t()) / (a.norm(dim=1)[:, None] * b.norm(dim=1)[None, :]))).float()
    if csls_k > 0:
        r1 = s.topk(csls_k, dim=1)[0].mean(1)
        r2 = s.topk(csls_k, dim=0)[0].mean(0)
        s = (2 * s.t() - r1).t() - r2
    order = torch.argsort(s, dim=1, descending=True)
    gold = torch.arange(s.shape[0], device=s.device)
    ranks = (order == gold[:, None]).int().argmax(1)
    n = s.shape[0]
    hits = [float((ranks < k).sum()) / n * 100 for k in top_k]
    return hits, float((ranks + 1).sum()) / n, float((1.0 / (ranks + 1).double()).sum()) / n, ranks


def get_neg(ILL, emb_src: torch.Tensor, emb_dst: torch.Tensor, k: int) -> torch.Tensor:
    """get_neg (modules/utils/util.py:31-54): top-k most similar entities of emb_dst per seed."""
    ILL = torch.as_tensor(np.asarray(ILL)).long().to(emb_src.device)
    s = emb_src[ILL] @ emb_dst.t()
    return s.topk(k, dim=1)[1].reshape(len(ILL) * k)


def compute_alignment_quality(e1: torch.Tensor, e2: torch.Tensor, list1, list2, scale=20.0):
    """compute_alignment_quality (train.py:231-259).  Returns (entropy, softmax_simi,
    softmax_simi2) exactly as the reference (dense N1xN2 -- small cases only)."""
    l1 = torch.as_tensor(np.asarray(list1)).long().to(e1.device)
    l2 = torch.as_tensor(np.asarray(list2)).long().to(e1.device)
    s = e1[l1] @ e2[l2].t()
    p1 = torch.softmax(s * scale, dim=1)
    h1 = (-torch.log(p1) * p1).sum(1).mean()
    p2 = torch.softmax(s.t() * scale, dim=1)
    h2 = (-torch.log(p2) * p2).sum(1).mean()
    full = e1 @ e2.t()
    m1 = torch.ones(e1.shape[0], dtype=torch.bool, device=e1.device)
    m1[l1] = False
    m2 = torch.ones(e2.shape[0], dtype=torch.bool, device=e1.device)
    m2[l2] = False
    full[m1] = -1
    full[:, m2] = -1
    return h1 + h2, torch.softmax(full * scale, dim=1), torch.softmax(full.t() * scale, dim=1)
