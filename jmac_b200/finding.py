"""Drop-ins for the reference's alignment scorer: ``modules/finding/similarity.py`` (sim, csls_sim),
``modules/finding/alignment.py`` (greedy_alignment), ``modules/finding/evaluation.py`` (valid, test),
``modules/utils/util.py`` (get_neg) and ``train.py:105-113`` (test_alignment_).

Same call signatures and return values.  The evaluation path never materialises the n1 x n2
similarity matrix: embeddings are rounded to bf16 once (rows L2-normalised first for the cosine
metric), and three passes of the tcgen05 similarity GEMM with a streaming epilogue produce
(A) row top-k means r1 and the gold-column similarity, (B) column top-k means r2 (transposed
product), (C) the rank of the gold column under CSLS = 2S - r1[i] - r2[j] by counting.
Ranks equal the reference's full-argsort ranks wherever the gold score differs from every
competitor's by more than bf16-input / accumulation-order noise (documented ties).
"""
from __future__ import annotations

import logging
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .functional import as_index

_DEF_DEV = "cuda"


def _dev_f32(x, device=None) -> torch.Tensor:
    t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    if not torch.cuda.is_available():
        raise _lib.JmacError("jmac_b200.finding needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device or (t.device if t.is_cuda else _DEF_DEV))
    return t.detach().to(dev, torch.float32, non_blocking=True).contiguous()


def to_bf16_rows(x: torch.Tensor, idx=None, normalize=False) -> torch.Tensor:
    """[n, dpad] bf16 (dpad = d rounded up to 64), optionally gathering rows ``idx`` and L2-normalising."""
    lib = _lib.load()
    n_src, d = x.shape
    dpad = (d + 63) // 64 * 64
    if dpad > 320:
        raise NotImplementedError("similarity GEMM keeps the A block resident: d <= 320 in this version")
    if idx is not None:
        idx = as_index(idx, x.device).reshape(-1)
    n = int(idx.numel()) if idx is not None else n_src
    out = torch.empty(n, dpad, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.jmac_to_bf16_rows(ptr(x), d, ptr(idx), n, int(bool(normalize)), ptr(out), dpad, stream_ptr()), "jmac_to_bf16_rows")
    return out


def sim_topk(A: torch.Tensor, B: torch.Tensor, k: int, alpha: float = 1.0, r1=None, r2=None, gold_col=None, gold_val=None,
             want_count=False, want_diag=False):
    """One pass of the streaming similarity kernel.  Returns dict(topv, topi, count, diag)."""
    lib = _lib.load()
    n1, n2, dpad = A.shape[0], B.shape[0], A.shape[1]
    dev = A.device
    topv = torch.empty(n1, max(k, 1), dtype=torch.float32, device=dev) if k > 0 else None
    topi = torch.empty(n1, max(k, 1), dtype=torch.int32, device=dev) if k > 0 else None
    cnt = torch.empty(n1, dtype=torch.int32, device=dev) if want_count else None
    diag = torch.empty(n1, dtype=torch.float32, device=dev) if want_diag else None
    if r2 is not None:
        n2p = (n2 + 127) // 128 * 128            # the kernel stages r2 in 128-column tiles with a bulk copy
        if r2.numel() < n2p:
            r2 = torch.cat([r2.reshape(-1).float(), torch.zeros(n2p - r2.numel(), dtype=torch.float32, device=dev)])
        r2 = r2.contiguous()
    with torch.cuda.device(dev):
        check(lib.jmac_sim_topk(ptr(A), ptr(B), n1, n2, dpad, int(k), float(alpha), ptr(r1), ptr(r2), ptr(gold_col), ptr(gold_val), ptr(cnt),
                                ptr(diag), ptr(topv), ptr(topi), stream_ptr()), "jmac_sim_topk")
    return {"topv": topv, "topi": topi, "count": cnt, "diag": diag}


def csls_gold(alpha, diag, r1, r2, gold_col):
    lib = _lib.load()
    out = torch.empty_like(diag)
    with torch.cuda.device(diag.device):
        check(lib.jmac_csls_gold(float(alpha), ptr(diag), ptr(r1), ptr(r2), ptr(gold_col), int(diag.numel()), ptr(out), stream_ptr()), "jmac_csls_gold")
    return out


def alignment_ranks(embed1, embed2, metric="inner", normalize=False, csls_k=0, device=None):
    """0-based rank of the gold column (row i <-> column i) for every row, plus the top-1 column.
    The streaming equivalent of sim() + csls_sim() + calculate_rank(accurate=True)."""
    e1, e2 = _dev_f32(embed1, device), _dev_f32(embed2, device)
    if metric not in ("inner", "cosine"):
        raise NotImplementedError(f"metric {metric!r}: the reference's hot path uses 'cosine' (train.py:94) and 'inner'")
    norm = bool(normalize) or metric == "cosine"
    A, B = to_bf16_rows(e1, normalize=norm), to_bf16_rows(e2, normalize=norm)
    n1, n2 = A.shape[0], B.shape[0]
    gold_col = torch.arange(n1, dtype=torch.int32, device=A.device)
    if csls_k > 0:
        pa = sim_topk(A, B, csls_k, gold_col=gold_col, want_diag=True)
        pb = sim_topk(B, A, csls_k)
        r1 = pa["topv"].mean(1).contiguous()                      # calculate_nearest_k: mean of the k largest per row
        r2 = pb["topv"].mean(1).contiguous()
        alpha = 2.0
    else:
        pa = sim_topk(A, B, 0, gold_col=gold_col, want_diag=True)
        r1 = r2 = None
        alpha = 1.0
    gold_val = csls_gold(alpha, pa["diag"], r1, r2, gold_col if r2 is not None else None)
    pc = sim_topk(A, B, 1, alpha=alpha, r1=r1, r2=r2, gold_val=gold_val, want_count=True)
    return pc["count"], pc["topi"][:, 0]


def alignment_ranks_bf16(A, B, csls_k=10, group=None):
    """Same three passes on already-converted bf16 tables ([n, dpad], gold column of row i = i).
    With ``group`` the work is ROW-SHARDED by source entity across the ranks (every rank holds both
    tables, 512 MB each at 1M x 256): pass A and C on this rank's rows of A, pass B on this rank's rows
    of B, then an all-gather of r2 and a final all-gather of the per-row results (rows are disjoint,
    so the "top-k merge" is a concatenation).  Returns (ranks[n1], top1[n1]) on every rank."""
    import torch.distributed as dist
    n1, n2 = A.shape[0], B.shape[0]
    dev = A.device
    world = dist.get_world_size(group) if group is not None else 1
    rank = dist.get_rank(group) if group is not None else 0

    def span(n):
        per = (n + world - 1) // world
        return min(rank * per, n), min((rank + 1) * per, n), per

    a_lo, a_hi, a_per = span(n1)
    b_lo, b_hi, b_per = span(n2)
    A_loc, B_loc = A[a_lo:a_hi], B[b_lo:b_hi]
    gold_col = torch.arange(a_lo, a_hi, dtype=torch.int32, device=dev)
    if csls_k > 0:
        pa = sim_topk(A_loc, B, csls_k, gold_col=gold_col, want_diag=True)
        pb = sim_topk(B_loc, A, csls_k)
        r1 = pa["topv"].mean(1).contiguous()
        r2_loc = pb["topv"].mean(1).contiguous()
        if world > 1:
            buf = torch.zeros(b_per, dtype=torch.float32, device=dev)
            buf[: r2_loc.numel()] = r2_loc
            r2_all = torch.empty(world * b_per, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(r2_all, buf, group=group)
            r2 = r2_all[:n2].contiguous()
        else:
            r2 = r2_loc
        alpha = 2.0
    else:
        pa = sim_topk(A_loc, B, 0, gold_col=gold_col, want_diag=True)
        r1 = r2 = None
        alpha = 1.0
    gold_val = csls_gold(alpha, pa["diag"], r1, r2, gold_col if r2 is not None else None)
    pc = sim_topk(A_loc, B, 1, alpha=alpha, r1=r1, r2=r2, gold_val=gold_val, want_count=True)
    ranks, top1 = pc["count"], pc["topi"][:, 0].contiguous()
    if world > 1:
        pack = torch.zeros(a_per, 2, dtype=torch.int32, device=dev)
        pack[: ranks.numel(), 0] = ranks
        pack[: ranks.numel(), 1] = top1
        allp = torch.empty(world * a_per, 2, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allp, pack, group=group)
        ranks, top1 = allp[:n1, 0].contiguous(), allp[:n1, 1].contiguous()
    return ranks, top1


def greedy_alignment(embed1, embed2, top_k, nums_threads, metric, normalize, csls_k, accurate, simi_matrix=None, return_simi=False,
                     logger=None):
    """modules/finding/alignment.py:10-69 -- returns (top_k, hits% rounded to 3 dp, mr, mrr)."""
    if simi_matrix is not None:
        raise NotImplementedError("a precomputed similarity matrix bypasses the streaming kernel; pass embeddings")
    assert 1 in top_k                                              # alignment.py:91
    ranks, _ = alignment_ranks(embed1, embed2, metric=metric, normalize=normalize, csls_k=csls_k)
    ranks = ranks.cpu().numpy().astype(np.int64)
    num = ranks.shape[0]
    hits = np.array([np.sum(ranks < k) for k in top_k]) / num * 100
    for i in range(len(hits)):
        hits[i] = round(hits[i], 3)
    mr = float(np.sum(ranks + 1) / num)
    mrr = float(np.sum(1.0 / (ranks + 1)) / num)
    if logger is not None:
        logger.info("Hits@{} = {}%, mr = {:.3f}, mrr = {:.6f}".format(top_k, hits, mr, mrr))
    return top_k, hits, mr, mrr


def valid(embeds1, embeds2, mapping, top_k, threads_num, metric='inner', normalize=False, csls_k=0, accurate=False, logger=None):
    """modules/finding/evaluation.py:9-17."""
    if mapping is not None:
        embeds1 = np.matmul(embeds1, mapping)
    _, hits1_12, mr_12, mrr_12 = greedy_alignment(embeds1, embeds2, top_k, threads_num, metric, normalize, csls_k, accurate, logger=logger)
    return hits1_12, mrr_12


def test(embeds1, embeds2, mapping, top_k, threads_num, metric='inner', normalize=False, csls_k=0, accurate=True, simi_matrix=None,
         return_simi=False, logger=None):
    """modules/finding/evaluation.py:20-28."""
    if mapping is not None:
        embeds1 = np.matmul(embeds1, mapping)
    return greedy_alignment(embeds1, embeds2, top_k, threads_num, metric, normalize, csls_k, accurate, simi_matrix=simi_matrix, logger=logger)


def test_alignment_(embeds11, embeds22, test_ent1, test_ent2, args, logger):
    """train.py:105-113."""
    embeds1 = np.array([embeds11[e] for e in test_ent1]) if not torch.is_tensor(embeds11) else embeds11[as_index(test_ent1, embeds11.device)]
    embeds2 = np.array([embeds22[e] for e in test_ent2]) if not torch.is_tensor(embeds22) else embeds22[as_index(test_ent2, embeds22.device)]
    return test(embeds1, embeds2, None, args.top_k, 8, metric=args.eval_metric, normalize=args.eval_norm, csls_k=args.csls, accurate=True,
                logger=logger)


def get_neg(ILL, emb_src, emb_dst, k):
    """modules/utils/util.py:31-54 -- indices of the k most similar rows of emb_dst for each seed,
    flattened to [len(ILL) * k] (LongTensor, on the CPU like the reference's inputs when they are)."""
    src, dst = _dev_f32(emb_src), _dev_f32(emb_dst)
    A = to_bf16_rows(src, idx=ILL)
    B = to_bf16_rows(dst)
    res = sim_topk(A, B, int(k))
    out = res["topi"].to(torch.int64).reshape(-1)
    if torch.is_tensor(emb_src) and not emb_src.is_cuda:
        out = out.cpu()
    return out


# ---- alignment quality / seed enlargement (train.py:231-259, 163-169) ---------------------------
def softmax_stats(A, B, alpha, row_mask=None, col_mask=None, masked_value=None):
    """Per-row softmax statistics of alpha * A.B^T (see jmac_sim_softmax_stats): dict(max, arg, Z, W)."""
    lib = _lib.load()
    n1, n2, dpad = A.shape[0], B.shape[0], A.shape[1]
    dev = A.device
    if col_mask is not None:
        n2p = (n2 + 127) // 128 * 128
        cm = torch.zeros(n2p, dtype=torch.float32, device=dev)
        cm[:n2] = col_mask.to(dev, torch.float32)
        col_mask = cm
    if row_mask is not None:
        row_mask = row_mask.to(dev, torch.float32).contiguous()
    out = {"max": torch.empty(n1, dtype=torch.float32, device=dev), "arg": torch.empty(n1, dtype=torch.int32, device=dev),
           "Z": torch.empty(n1, dtype=torch.float32, device=dev), "W": torch.empty(n1, dtype=torch.float32, device=dev)}
    with torch.cuda.device(dev):
        check(lib.jmac_sim_softmax_stats(ptr(A), ptr(B), n1, n2, dpad, float(alpha), ptr(row_mask), ptr(col_mask),
                                         float(-alpha if masked_value is None else masked_value), ptr(out["max"]), ptr(out["arg"]),
                                         ptr(out["Z"]), ptr(out["W"]), stream_ptr()), "jmac_sim_softmax_stats")
    return out


class SoftmaxRows:
    """What train.py:165-168 needs from the dense ``softmax_simi`` matrix, without the matrix:
    ``simi.max(dim=1)`` -> (row max probability, row argmax) and ``simi[rows].max(dim=1)`` likewise.
    (``torch.multinomial`` on the row-max vector stays in torch for RNG parity.)"""

    def __init__(self, maxp: torch.Tensor, arg: torch.Tensor, shape):
        self._maxp, self._arg, self.shape = maxp, arg, tuple(shape)

    def max(self, dim=1):
        if dim not in (1, -1):
            raise NotImplementedError("only row-wise max is available without materialising the matrix")
        return self._maxp, self._arg

    def __getitem__(self, rows):
        idx = as_index(rows, self._maxp.device) if not isinstance(rows, slice) else rows
        return SoftmaxRows(self._maxp[idx], self._arg[idx], (self._maxp[idx].shape[0], self.shape[1]))

    def __len__(self):
        return self.shape[0]


def compute_alignment_quality(embedding1, embedding2, list1, list2, scale=20.0):
    """train.py:231-259.  Returns (entropy, softmax_simi, softmax_simi2) where the two "matrices" are
    ``SoftmaxRows`` views (row max / argmax only) -- exactly what seed enlargement consumes
    (train.py:165-168) -- and ``entropy`` is a 0-dim CPU tensor as in the reference.
    Streaming: four passes of the similarity kernel in softmax-statistics mode; the reference's
    O(N*L) Python ``not in list`` scans and dense N1 x N2 CPU softmaxes are gone."""
    e1, e2 = _dev_f32(embedding1), _dev_f32(embedding2)
    dev = e1.device
    n1, n2 = e1.shape[0], e2.shape[0]
    l1, l2 = as_index(list1, dev), as_index(list2, dev)
    # (1) entropy over the sub-matrix of the valid alignment entities, both directions
    A_sub, B_sub = to_bf16_rows(e1, idx=l1), to_bf16_rows(e2, idx=l2)
    s12 = softmax_stats(A_sub, B_sub, scale)
    s21 = softmax_stats(B_sub, A_sub, scale)
    ent = (torch.log(s12["Z"]) - s12["W"] / s12["Z"]).mean() + (torch.log(s21["Z"]) - s21["W"] / s21["Z"]).mean()
    # (2) full matrix with rows not in list1 / columns not in list2 set to -1 (train.py:252-255)
    A, B = to_bf16_rows(e1), to_bf16_rows(e2)
    m1 = torch.ones(n1, dtype=torch.float32, device=dev)
    m1[l1] = 0
    m2 = torch.ones(n2, dtype=torch.float32, device=dev)
    m2[l2] = 0
    f12 = softmax_stats(A, B, scale, row_mask=m1, col_mask=m2)
    f21 = softmax_stats(B, A, scale, row_mask=m2, col_mask=m1)
    out_cpu = torch.is_tensor(embedding1) and not embedding1.is_cuda

    def view(f, shape):
        maxp, arg = 1.0 / f["Z"], f["arg"].to(torch.int64)
        if out_cpu:
            maxp, arg = maxp.cpu(), arg.cpu()
        return SoftmaxRows(maxp, arg, shape)

    ent = ent.cpu() if out_cpu else ent
    return ent, view(f12, (n1, n2)), view(f21, (n2, n1))


# ---- dense convenience API (small inputs only; not the hot path) ---------------------------------
def sim(embed1, embed2, metric='inner', normalize=False, csls_k=0):
    """modules/finding/similarity.py:13-55 for callers that really want the dense matrix (numpy
    float32 [n1,n2]).  Library fp32 matmul; the evaluation path uses alignment_ranks() instead."""
    e1, e2 = _dev_f32(embed1), _dev_f32(embed2)
    if normalize or metric == "cosine":
        e1 = torch.nn.functional.normalize(e1, 2, -1)
        e2 = torch.nn.functional.normalize(e2, 2, -1)
    elif metric != "inner":
        raise NotImplementedError(metric)
    s = e1 @ e2.t()
    if csls_k > 0:
        s = _csls_dense(s, csls_k)
    return s.cpu().numpy()


def _csls_dense(s: torch.Tensor, k: int) -> torch.Tensor:
    r1 = s.topk(k, dim=1)[0].mean(1)
    r2 = s.topk(k, dim=0)[0].mean(0)
    return 2 * s - r1[:, None] - r2[None, :]


def csls_sim(sim_mat, k):
    """modules/finding/similarity.py:58-78 on a dense matrix."""
    return _csls_dense(_dev_f32(sim_mat), k).cpu().numpy()
