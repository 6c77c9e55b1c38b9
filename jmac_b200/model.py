"""Drop-in replacements for ``RelationAwareLayer`` and ``JMAC`` of the reference's
``src/jmac_model.py`` -- same constructor arguments, parameter / ``state_dict`` names, method
signatures, return values and error behaviour -- with the hot path running in hand-written sm_100a
kernels (libjmac_b200.so) instead of un-fused eager ops.  Host code stays PyTorch.

What runs where
  * per-edge work (gathers, composition, attention logits, edge softmax, sqrt(deg) normalisation,
    scatter-sum, and all of its backward), BatchNorm+tanh, the TransE-L1 margin scorer, the pair
    cosine of the alignment losses and the L1 link-prediction distances: CUDA kernels;
  * the node / relation-level dense products (X W_top, X W_bot, R' W_bot, (.) gcn_weight, the
    relation MLPs): ``torch.mm`` (cuBLAS fp32, plain library GEMMs) with autograd.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.init import xavier_normal_

from . import functional as JF


def mult_block_bounds(in_degree: torch.Tensor, e_max: int) -> List[int]:
    """Boundaries 0 = b[0] < b[1] < ... = N of consecutive destination blocks with at most ``e_max`` in-edges each; a node with more
    in-edges than that is a block of its own (comp_op='mult': the per-edge tables are built one block at a time)."""
    n = int(in_degree.numel())
    cum = torch.cumsum(in_degree.to(torch.int64), 0)
    bounds, done = [0], 0
    while bounds[-1] < n:
        b = int(torch.searchsorted(cum, torch.tensor(done + int(e_max), dtype=torch.int64), right=True))   # first node whose cumulative count exceeds the budget
        b = min(max(b, bounds[-1] + 1), n)
        bounds.append(b)
        done = int(cum[b - 1])
    return bounds


def get_param(shape):
    """modules/helper/helper.py:68-74 -- xavier_normal_ initialised Parameter."""
    param = nn.Parameter(torch.empty(*shape))
    xavier_normal_(param.data)
    return param


class RelationAwareLayer(nn.Module):
    """src/jmac_model.py:10-109.  ``forward(ent_emb, rel_emb, edge_index, edge_type) -> [N, out]``.

    The reference computes, per edge, ``cat(x_i, x_j - r_t) @ w_att`` and ``(x_j - r_t) @ gcn_weight``
    (two E-row GEMMs).  Here both are hoisted to node level: with
    ``w_att = [W_top; W_bot]``, ``P = X W_top``, ``Q = X W_bot``, ``S = R' W_bot`` the logit is
    ``a . lrelu(P[i] + Q[j] - S[t])`` and the message GEMM is applied once after aggregation.
    The self-loop ``propagate`` (:50) has a singleton softmax (== 1) and no norm, so it reduces to
    ``(X - R'[loop]) @ gcn_weight``; attention parameters get no gradient from it, as in the reference.
    """

    def __init__(self, in_channels, out_channels, rel_dim, act=lambda x: x, args=None):
        super().__init__()
        if in_channels != out_channels:
            raise NotImplementedError("the reference only instantiates square layers (src/jmac_model.py:147-149)")
        self.layer_act = act
        self.args = args
        self.rel_transform_weight1 = get_param((rel_dim, in_channels))
        self.rel_transform_weight2 = get_param((in_channels, in_channels))
        self.gcn_weight = get_param((in_channels, out_channels))
        self.loop_rel = get_param((1, rel_dim))
        self.w_att = get_param((2 * in_channels, out_channels))
        self.a_att = get_param((out_channels, 1))
        self.leaky_slope = float(args.leaky_relu_w)
        self.rel_act = getattr(args, "rel_act", "leaky")          # 'relu' = JMAC_DBPv1/models/jmac_model.py:27,51
        self.bn = nn.BatchNorm1d(out_channels)
        self.comp_op = args.comp_op
        self.in_channels = in_channels
        self.dist_group = None      # set by jmac_b200.dist for destination-sharded execution

    def relation_transform(self, rel_emb):
        rel = torch.cat([rel_emb, self.loop_rel], dim=0)                        # :39
        rel = JF.node_mm(rel, self.rel_transform_weight1)                       # :40  (tensor-core product from 512 rows up, library below)
        rel = F.leaky_relu(rel, self.leaky_slope) if self.rel_act == "leaky" else F.relu(rel)   # :41
        return JF.node_mm(rel, self.rel_transform_weight2)                      # :42

    def forward(self, ent_emb, rel_emb, edge_index, edge_type):
        if self.comp_op == "mult":
            return self.forward_mult(ent_emb, rel_emb, edge_index, edge_type)
        if self.comp_op != "sub":
            raise NotImplementedError                                            # :65-66
        if self.layer_act is not torch.tanh:
            raise NotImplementedError("the fused BN+activation kernel implements tanh (BaseModel.act, src/jmac_model.py:118)")
        num_ent = ent_emb.size(0)
        d = self.in_channels
        Rp = self.relation_transform(rel_emb)                                    # [nr+1, d]; row nr = self loop
        plan = JF.get_plan(edge_index, edge_type, num_ent, Rp.size(0))
        X = ent_emb.contiguous()
        w_att, a_att, gcn = self.w_att, self.a_att, self.gcn_weight
        bn_w, bn_b, run_m, run_v = self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var
        # Widths that are not a multiple of 32 (JMAC_DBPv1: d = 300, JMAC_DBPv1/main_with_args.py:36) run at the next multiple of 32 so
        # that the node products stay on the tensor cores: zero columns / weight rows / BN affine entries are appended (autograd slices
        # the gradients back), the padded columns stay exactly zero through the layer, and the result is cut back to d columns.
        dp = JF.padded_width(d)
        pad = dp - d if (dp != d and dp <= 512 and X.is_cuda and X.dtype == torch.float32 and num_ent >= JF._TC_MIN_ROWS
                         and self.dist_group is None) else 0
        if pad:
            X, Rp = F.pad(X, (0, pad)), F.pad(Rp, (0, pad))
            w_att = torch.cat([F.pad(w_att[:d], (0, pad, 0, pad)), F.pad(w_att[d:], (0, pad, 0, pad))], dim=0)
            a_att, gcn = F.pad(a_att, (0, 0, 0, pad)), F.pad(gcn, (0, pad, 0, pad))
            bn_w = F.pad(bn_w, (0, pad)) if bn_w is not None else None
            bn_b = F.pad(bn_b, (0, pad)) if bn_b is not None else None
            run_m = F.pad(run_m, (0, pad)) if run_m is not None else None                 # temporaries: the kernel updates them in place
            run_v = F.pad(run_v, (0, pad), value=1.0) if run_v is not None else None
        dw = d + pad
        if JF.attn_edge_eligible(X, dw):
            # large node tables: one autograd node for the attention inputs + edge phase, so that the input gradient
            # dP W_top^T + dQ W_bot^T + dXe + dT is accumulated in GEMM epilogues instead of elementwise passes
            T = JF.attn_edge(plan, X, w_att, Rp, a_att, self.leaky_slope)
        else:
            P = JF.node_mm(X, w_att[:dw])           # node tables: 3xTF32 tcgen05 GEMM (fp32-accurate), small ones: library
            Q = JF.node_mm(X, w_att[dw:])
            S = torch.mm(Rp, w_att[dw:])
            T = JF.edge_phase(plan, P, Q, X, S, Rp, a_att, self.leaky_slope, X)     # agg + X - R'[loop] (self-loop fused)
        pre = JF.node_mm(T, gcn * 0.5)                                           # (neigh + self)/2, :49-52
        training = self.training or not self.bn.track_running_stats
        if self.training and self.bn.track_running_stats:
            self.bn.num_batches_tracked += 1
        momentum = self.bn.momentum if self.bn.momentum is not None else 0.1
        out = JF.bn_tanh(pre, bn_w, bn_b, run_m, run_v, momentum, self.bn.eps, training, self.dist_group)
        if pad:
            if self.training and self.bn.track_running_stats and run_m is not None:
                with torch.no_grad():
                    self.bn.running_mean.copy_(run_m[:d])
                    self.bn.running_var.copy_(run_v[:d])
            out = out[:, :d].contiguous()
        return out


    def forward_mult(self, ent_emb, rel_emb, edge_index, edge_type):
        """comp_op='mult' (src/jmac_model.py:63-64): the composed row ``m_e = x_j * r_t`` depends on the edge, so the
        attention product ``m_e W_bot`` cannot be hoisted to the nodes.  The per-edge table ``M`` is materialised (gather +
        multiply), its attention projection is ONE product on the tcgen05 3xTF32 kernel, and the edge kernels run unchanged
        on a bipartite plan whose "sources" are the rows of M (zero relation tables): logit ``a . lrelu(P[i] + (M W_bot)[e])``,
        aggregate ``sqrt(deg_i) sum_e alpha_e M[e]``; the message GEMM is still applied once per node after aggregation.  The
        self loop is ``(x_i * r_loop) W_g``.  Beyond ``MULT_TABLE_BYTES`` the tables are built one block of destinations at a
        time and recomputed in backward (``mult_edge_phase``): memory is O(block edges x d), never O(E d)."""
        if self.layer_act is not torch.tanh:
            raise NotImplementedError("the fused BN+activation kernel implements tanh (BaseModel.act, src/jmac_model.py:118)")
        num_ent = ent_emb.size(0)
        Rp = self.relation_transform(rel_emb)
        X = ent_emb.contiguous()
        ei = JF.as_index(edge_index, X.device)
        et = JF.as_index(edge_type, X.device)
        T = self.mult_edge_phase(X, X, Rp, ei, et)
        pre = JF.node_mm(T, self.gcn_weight * 0.5)
        training = self.training or not self.bn.track_running_stats
        if self.training and self.bn.track_running_stats:
            self.bn.num_batches_tracked += 1
        momentum = self.bn.momentum if self.bn.momentum is not None else 0.1
        return JF.bn_tanh(pre, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var, momentum, self.bn.eps,
                          training, None)

    # per-edge tables of comp_op='mult' larger than this are built per block of destinations (bytes of ONE [E_block, d] fp32 table)
    MULT_TABLE_BYTES = 1 << 30

    def _mult_blocks(self, ei, et, n_dst, d, ops=JF):
        """Destination blocks [lo, hi) whose in-edges fit ``MULT_TABLE_BYTES`` (a hub with more edges than that is a block of its own),
        each with its edge selection, source rows, types and edge-table plan.  Cached per edge list."""
        key = (ei.data_ptr(), tuple(ei.shape), ei._version, et.data_ptr(), et._version, int(n_dst), int(d), int(self.MULT_TABLE_BYTES), id(ops))
        cache = self.__dict__.setdefault("_mult_block_cache", {})
        if key in cache:
            return cache[key][0]
        E = int(ei.shape[1])
        e_max = max(1, int(self.MULT_TABLE_BYTES) // (4 * d))
        row, col = ei[0], ei[1]
        if E <= e_max:
            if ops is JF:
                plan = JF.get_edge_table_plan(ei, n_dst)
            else:
                plan = ops.GraphPlan(torch.stack([row, torch.arange(E, device=ei.device)]), torch.zeros(E, dtype=torch.int64, device=ei.device), n_dst, 1,
                                     n_src=max(E, 1))
            blocks = [(0, n_dst, None, col, et, plan)]
        else:
            bounds = mult_block_bounds(torch.bincount(row, minlength=n_dst).cpu(), e_max)
            blocks = []
            for lo, hi in zip(bounds, bounds[1:]):
                sel = torch.nonzero((row >= lo) & (row < hi)).reshape(-1)
                ei_b = torch.stack([row[sel] - lo, torch.arange(sel.numel(), device=ei.device)])
                plan = ops.GraphPlan(ei_b, torch.zeros(sel.numel(), dtype=torch.int64, device=ei.device), hi - lo, 1, n_src=max(int(sel.numel()), 1))
                blocks.append((lo, hi, sel, col[sel], et[sel], plan))
        cache.clear()                                             # one edge list at a time
        cache[key] = (blocks, ei, et)
        return blocks

    def mult_edge_phase(self, X_dst, X_src, Rp, ei, et, ops=None):
        """T = sqrt(deg_i) sum_e alpha_e (x_j * r_t) + x_i * r_loop for the destinations ``X_dst`` (rows ei[0]) with sources
        ``X_src`` (rows ei[1]; the same table on one GPU, the gathered table under destination sharding).  ``ops``: the compute
        backend (``jmac_b200.functional``; the gloo tests inject torch stand-ins)."""
        ops = ops or JF
        node_mm = getattr(ops, "node_mm", torch.mm)
        d = self.in_channels
        n_dst = X_dst.size(0)
        Wt, Wb = self.w_att[:d], self.w_att[d:]
        P = node_mm(X_dst, Wt)
        blocks = self._mult_blocks(ei, et, n_dst, d, ops)
        slope = self.leaky_slope

        def run(plan, col_b, et_b, P_b, Xd_b, X_src, Rp, Wb, a_att):
            if col_b.numel():
                M = X_src.index_select(0, col_b) * Rp.index_select(0, et_b)
            else:
                M = X_src.new_zeros((1, d))
            Qm = node_mm(M, Wb)
            zero = X_src.new_zeros((1, d))
            return ops.edge_phase(plan, P_b, Qm, M, zero, zero, a_att, slope, Xd_b * Rp[-1])

        if len(blocks) == 1:
            lo, hi, _, col_b, et_b, plan = blocks[0]
            return run(plan, col_b, et_b, P, X_dst, X_src, Rp, Wb, self.a_att)
        from torch.utils.checkpoint import checkpoint
        outs = []
        for lo, hi, _, col_b, et_b, plan in blocks:
            # the block's [E_block, d] tables live only inside this call, forward and (recomputed) backward
            outs.append(checkpoint(run, plan, col_b, et_b, P[lo:hi], X_dst[lo:hi], X_src, Rp, Wb, self.a_att, use_reentrant=False))
        return torch.cat(outs, dim=0)


class BaseModel(nn.Module):
    """src/jmac_model.py:112-122."""

    def __init__(self):
        super().__init__()
        self.act = torch.tanh
        self.bceloss = nn.BCELoss()

    def triple_loss(self, pred, true_label):
        return self.bceloss(pred, true_label)


class JMAC(BaseModel):
    """src/jmac_model.py:125-380 -- same parameters (names, shapes, init order) and methods."""

    def __init__(self, args, entity_name_emb, num_relations, num_entities):
        super().__init__()
        self.args = args
        self.ent_info_att = torch.FloatTensor(entity_name_emb)
        assert entity_name_emb.shape[0] == num_entities
        self.entity_dim = args.dim
        self.relation_dim = args.dim
        self.device = args.device
        d = self.entity_dim
        self.ent_init_att_completion = get_param((num_entities, d))
        self.rel_init_att_completion = get_param((num_relations, d))
        self.rel_init_att_alignment = get_param((num_relations, d))

        self.completion_dropout = nn.Dropout(args.dropout)
        self.atv_mlp = nn.LeakyReLU(args.leaky_relu_w)
        self.conv1_alignment = RelationAwareLayer(d, d, rel_dim=d, act=self.act, args=args)
        self.conv2_alignment = RelationAwareLayer(d, d, rel_dim=d, act=self.act, args=args)
        self.conv1_completion = RelationAwareLayer(d, d, rel_dim=d, act=self.act, args=args)

        self.name_linear = get_param((self.ent_info_att.shape[1], d))
        self.margin_align = args.margin_align
        self.k = args.num_negative
        self.margin_completion = nn.Parameter(torch.Tensor([args.margin_completion]))
        self.margin_completion.requires_grad = False

        L = args.num_gcn_layer
        self.uni_linear1_1, self.uni_linear1_2 = get_param((d * L, d)), get_param((d, d))
        self.uni_linear2_1, self.uni_linear2_2 = get_param((d * L, d)), get_param((d, d))
        self.rel_linear11, self.rel_linear12 = get_param((d, d)), get_param((d, d))
        self.rel_linear21, self.rel_linear22 = get_param((d, d)), get_param((d, d))
        self.rel_linear11_uni, self.rel_linear12_uni = get_param((d, d)), get_param((d, d))
        self.all_linear_completion = get_param((d * (L + 1), d))

        self.forward_base = self.forward_no_name if self.args.no_name_info else self.forward_name
        self._name_dev = None     # device-resident copy of ent_info_att (the reference re-uploads it every call, :177)

    # ---- encoders -------------------------------------------------------------------------
    def _name_table(self, device):
        if self._name_dev is None or self._name_dev.device != torch.device(device):
            self._name_dev = self.ent_info_att.to(device)
        return self._name_dev

    def forward_name(self, edge_index, edge_type, ent_bases, rel_bases):
        """:172-204."""
        lo, hi = ent_bases
        rlo, rhi = rel_bases
        dev = self.ent_init_att_completion.device
        e0 = self.ent_init_att_completion[lo:hi]
        name = torch.mm(self._name_table(dev)[lo:hi], self.name_linear)
        e0n = self.completion_dropout(JF.row_normalize(e0))
        a0 = JF.concat_linear((e0n, name), self.uni_linear1_1)            # cat(e0n, name) @ W without the concatenation
        a1 = self.conv1_alignment(a0, self.rel_init_att_alignment[rlo:rhi], edge_index, edge_type)
        align_embs = [a0, a1]
        comp_embs = [e0]
        comp_rels = [self.rel_init_att_completion[rlo:rhi]]
        if self.args.num_gcn_layer == 2:
            c1 = self.conv1_completion(e0, self.rel_init_att_completion[rlo:rhi], edge_index, edge_type)
            c1n = self.completion_dropout(JF.row_normalize(c1))
            a_in = JF.concat_linear((c1n, a1), self.uni_linear2_1)
            r1 = torch.mm(self.atv_mlp(torch.mm(self.rel_init_att_completion[rlo:rhi], self.rel_linear11)), self.rel_linear12)
            ar = torch.mm(self.atv_mlp(torch.mm(self.rel_init_att_alignment[rlo:rhi], self.rel_linear11_uni)), self.rel_linear12_uni)
            a2 = self.conv2_alignment(a_in, ar, edge_index, edge_type)
            align_embs.append(a2)
            comp_embs.append(c1)
            comp_rels.append(r1)
        out = JF.concat_linear(align_embs, self.all_linear_completion)
        return out, comp_embs, comp_rels

    def forward_no_name(self, edge_index, edge_type, ent_bases, rel_bases):
        """:207-220."""
        e0 = self.ent_init_att_completion[ent_bases[0]:ent_bases[1]]
        r0 = self.rel_init_att_completion[rel_bases[0]:rel_bases[1]]
        comp_embs, comp_rels = [e0], [r0]
        if self.args.num_gcn_layer == 2:
            e1 = self.conv1_completion(e0, r0, edge_index, edge_type)
            r1 = JF.node_mm(self.atv_mlp(JF.node_mm(r0, self.rel_linear11)), self.rel_linear12)
            comp_embs.append(e1)
            comp_rels.append(r1)
        return comp_embs[-1], comp_embs, comp_rels

    def get_emb(self, edge_index, edge_type, ent_bases, rel_bases, pyt=False):
        """:223-234."""
        a, comps, _ = self.forward_base(edge_index, edge_type, ent_bases, rel_bases)
        a = JF.row_normalize(a)
        c = JF.row_normalize(comps[-1])
        if pyt:
            return a.detach().cpu(), c.detach().cpu()
        return a.detach().cpu().numpy(), c.detach().cpu().numpy()

    # ---- losses ---------------------------------------------------------------------------
    def alignment_loss_simple(self, links, ent_embeddings1, ent_embeddings2):
        """:237-249 -- mean cosine distance of the seed links."""
        if not len(links):
            return 0
        links = np.asarray(links) if not torch.is_tensor(links) else links
        cosv = JF.pair_cos(ent_embeddings1, ent_embeddings2, links[:, 0].reshape(-1), links[:, 1].reshape(-1))
        return (1 - cosv).mean()

    def alignment_loss(self, feeddict, edge_index1, edge_type1, edge_index2, edge_type2):
        """:252-292."""
        links = feeddict["links"]
        if not len(links):
            return 0
        e1, _, _ = self.forward_base(edge_index1, edge_type1, feeddict["ent_bases1"], feeddict["rel_bases1"])
        e2, _, _ = self.forward_base(edge_index2, edge_type2, feeddict["ent_bases2"], feeddict["rel_bases2"])
        links = np.asarray(links) if not torch.is_tensor(links) else links
        n = len(links)
        A = 1 - JF.pair_cos(e1, e2, links[:, 0].reshape(-1), links[:, 1].reshape(-1))
        D = A + self.margin_align
        total = 0
        for left, right in (("neg_left", "neg_right"), ("neg2_left", "neg2_right")):
            Bc = 1 - JF.pair_cos(e1, e2, feeddict[left], feeddict[right])
            total = total + F.relu(-Bc.view(n, -1) + D.view(n, 1)).sum()
        return total / (2 * self.k * n) * 1

    def forward_linkpred(self, e_index, r_index, edge_index, edge_type, all_index, ent_bases, rel_bases, pred_head=False):
        """:295-313.  ``all_index`` is ``list(range(N))`` at every call site (src/validate.py:44); any other index list
        selects / permutes the table first, as the reference's ``ent[all_index]`` does (:304).
        The [B,N] distances come from ``jmac_l1_dist``; when a gradient is requested through them (the reference's result
        carries one, although no caller uses it) they are formed by the differentiable library ``torch.cdist`` instead."""
        _, ents, rels = self.forward_base(edge_index, edge_type, ent_bases, rel_bases)
        L = self.args.num_gcn_layer
        n = ents[0].shape[0]
        identity = isinstance(all_index, range) and all_index == range(n)
        if not identity:
            ai = np.asarray(all_index.cpu() if torch.is_tensor(all_index) else all_index)
            identity = ai.shape == (n,) and bool(np.array_equal(ai, np.arange(n)))
        if not identity:
            idx = JF.as_index(all_index, ents[0].device)
            ents = [e[idx] for e in ents]                  # all_kg_emb = ent[all_index] (:304); e_index addresses THIS table (:306)
        if torch.is_grad_enabled() and any(t.requires_grad for t in list(ents[:L]) + list(rels[:L])):
            ei, ri = JF.as_index(e_index, ents[0].device), JF.as_index(r_index, ents[0].device)
            dist = 0
            for e, r in zip(ents[:L], rels[:L]):
                er = e[ei] - r[ri] if pred_head else e[ei] + r[ri]
                dist = dist + torch.cdist(er, e, p=1)
            return dist
        return JF.l1_dist(ents[:L], rels[:L], e_index, r_index, pred_head)

    def completion_loss(self, data, edge_index1, edge_type1, edge_index2, edge_type2, feeddict, source=True):
        """:316-380."""
        return self.completion_loss_multi([data], edge_index1, edge_type1, edge_index2, edge_type2, feeddict, source)[0]

    def completion_loss_multi(self, batches, edge_index1, edge_type1, edge_index2, edge_type2, feeddict, source=True):
        """Optional multi-batch step: the reference re-encodes both KGs for every 1000-triple batch
        (:325-326, train.py:332-361); here both encoders run ONCE and every batch of ``batches`` (dicts with
        batch_h / batch_r / batch_t in the layout of train.py:347-352) is scored against the same embeddings.
        Returns one loss per batch, each equal to what ``completion_loss`` returns for that batch at the current
        parameters; ``sum(losses).backward()`` accumulates the gradient of all of them through one encoder backward."""
        links = feeddict["links"]
        _, ents1, rels1 = self.forward_base(edge_index1, edge_type1, feeddict["ent_bases1"], feeddict["rel_bases1"])
        _, ents2, rels2 = self.forward_base(edge_index2, edge_type2, feeddict["ent_bases2"], feeddict["rel_bases2"])
        L = self.args.num_gcn_layer
        ents, rels = (ents1, rels1) if source else (ents2, rels2)
        if getattr(self.args, "normalize_triples", False):   # JMAC_DBPv1/models/jmac_model.py:245-247 scores unit-norm h, r, t
            ents, rels = [JF.row_normalize(e) for e in ents[:L]], [JF.row_normalize(r) for r in rels[:L]]
        align = 0
        for layer in range(L):
            align = align + self.alignment_loss_simple(links, ents1[layer], ents2[layer])
        losses = []
        for data in batches:
            per_layer = JF.transe_margin_loss(list(ents[:L]), list(rels[:L]), data["batch_h"], data["batch_r"], data["batch_t"],
                                              self.args.batch_size, self._margin())
            losses.append(per_layer.sum().view(1) + align)      # shape [1] like the reference (margin Parameter broadcast, :374)
        return losses

    def _margin(self) -> float:
        # margin_completion is a frozen Parameter (:156-157); its value is cached to avoid a device sync per step and
        # dropped whenever the parameter object can have changed (load_state_dict, .to() / .cuda() / .double())
        m = getattr(self, "_margin_cache", None)
        if m is None:
            m = float(self.margin_completion.detach().cpu().item())
            self._margin_cache = m
        return m

    def _load_from_state_dict(self, *args, **kwargs):
        self._margin_cache = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._margin_cache = None
        return super()._apply(fn, *args, **kwargs)
