"""Drop-ins for the reference's alignment scorer: ``modules/finding/similarity.py`` (sim, csls_sim),
``modules/finding/alignment.py`` (greedy_alignment), ``modules/finding/evaluation.py`` (valid, test),
``modules/utils/util.py`` (get_neg) and ``train.py:105-113`` (test_alignment_).

Same call signatures and return values.  The evaluation path never materialises the n1 x n2
similarity matrix: three passes of the tcgen05 similarity GEMM with a streaming epilogue produce
(A) row top-k means r1 and the gold-column similarity, (B) column top-k means r2 (transposed
product), (C) the rank of the gold column under CSLS = 2S - r1[i] - r2[j] by counting.

Precision.  The reference evaluates in float64 / float32 (cosine by f64 ``cdist`` cast to f32,
similarity.py:44-46; ``get_neg`` and ``compute_alignment_quality`` by fp32 ``mm``, util.py:52,
train.py:239-257).  The default here is ``precision='fp32'``: the fp32-accurate 3xTF32 mode of the
kernel (|S - exact| ~ 1e-6, the accuracy of an fp32 GEMM), so ranks, top-k index sets and softmax
statistics equal the reference's wherever two compared scores differ by more than fp32 rounding
noise (documented ties).  ``precision='bf16'`` rounds the embeddings to bf16 first
(one MMA per K step instead of three): the mode for the 1M x 1M problem (``alignment_ranks_bf16``).
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
DEFAULT_PRECISION = "fp32"


def _dev_f32(x, device=None) -> torch.Tensor:
    t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    if not torch.cuda.is_available():
        raise _lib.JmacError("jmac_b200.finding needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device or (t.device if t.is_cuda else _DEF_DEV))
    return t.detach().to(dev, torch.float32, non_blocking=True).contiguous()


class SimTable:
    """Operand table of the similarity kernel: bf16 rows (``lo is None``) or the hi / lo fp32 pair of the
    3xTF32 product.  ``[n, dpad]``, zero-padded columns."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: Optional[torch.Tensor] = None):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    @property
    def device(self):
        return self.hi.device

    @property
    def f32(self) -> bool:
        return self.lo is not None

    def __getitem__(self, sl):
        if not isinstance(sl, slice):
            raise TypeError("SimTable supports row slices only")
        return SimTable(self.hi[sl], None if self.lo is None else self.lo[sl])


def _table(t) -> SimTable:
    return t if isinstance(t, SimTable) else SimTable(t)


def to_bf16_rows(x: torch.Tensor, idx=None, normalize=False) -> torch.Tensor:
    """[n, dpad] bf16 (dpad = d rounded up to 64), optionally gathering rows ``idx`` and L2-normalising."""
    lib = _lib.load()
    n_src, d = x.shape
    dpad = (d + 63) // 64 * 64
    if dpad > 320:
        raise NotImplementedError("the bf16 similarity GEMM keeps the A block resident: d <= 320 (use precision='fp32')")
    if idx is not None:
        idx = as_index(idx, x.device).reshape(-1)
    n = int(idx.numel()) if idx is not None else n_src
    out = torch.empty(n, dpad, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.jmac_to_bf16_rows(ptr(x), d, ptr(idx), n, int(bool(normalize)), ptr(out), dpad, stream_ptr()), "jmac_to_bf16_rows")
    return out


def to_tf32_rows(x: torch.Tensor, idx=None, normalize=False) -> SimTable:
    """The hi / lo pair of the 3xTF32 product, [n, dpad] fp32 each (dpad = d rounded up to 32)."""
    lib = _lib.load()
    n_src, d = x.shape
    dpad = (d + 31) // 32 * 32
    if idx is not None:
        idx = as_index(idx, x.device).reshape(-1)
    n = int(idx.numel()) if idx is not None else n_src
    hi = torch.empty(n, dpad, dtype=torch.float32, device=x.device)
    lo = torch.empty(n, dpad, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.jmac_to_tf32_rows(ptr(x), d, ptr(idx), n, int(bool(normalize)), ptr(hi), ptr(lo), dpad, stream_ptr()), "jmac_to_tf32_rows")
    return SimTable(hi, lo)


def to_rows(x: torch.Tensor, idx=None, normalize=False, precision: Optional[str] = None) -> SimTable:
    precision = precision or DEFAULT_PRECISION
    if precision == "fp32":
        return to_tf32_rows(x, idx, normalize)
    if precision == "bf16":
        return SimTable(to_bf16_rows(x, idx, normalize))
    raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")


def _workspace(lib, n1, n2, k, dev):
    nbytes = int(lib.jmac_sim_workspace_bytes(n1, n2, int(k)))
    return (torch.empty(nbytes, dtype=torch.uint8, device=dev), nbytes) if nbytes > 0 else (None, 0)


def sim_topk(A, B, k: int, alpha: float = 1.0, r1=None, r2=None, gold_col=None, gold_val=None, want_count=False, want_diag=False,
             split: bool = True):
    """One pass of the streaming similarity kernel.  ``A``, ``B``: SimTable (or raw bf16 tensors) of the same
    precision.  ``split=False`` withholds the workspace (one CTA per 256 rows; for tests: the results are
    bit-identical).  Returns dict(topv, topi, count, diag)."""
    lib = _lib.load()
    A, B = _table(A), _table(B)
    if A.f32 != B.f32:
        raise ValueError("both operands must have the same precision")
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
    ws, ws_bytes = _workspace(lib, n1, n2, k, dev) if split else (None, 0)
    with torch.cuda.device(dev):
        if A.f32:
            check(lib.jmac_sim_topk_f32(ptr(A.hi), ptr(A.lo), ptr(B.hi), ptr(B.lo), n1, n2, dpad, int(k), float(alpha), ptr(r1), ptr(r2),
                                        ptr(gold_col), ptr(gold_val), ptr(cnt), ptr(diag), ptr(topv), ptr(topi), ptr(ws), ws_bytes, stream_ptr()),
                  "jmac_sim_topk_f32")
        else:
            check(lib.jmac_sim_topk(ptr(A.hi), ptr(B.hi), n1, n2, dpad, int(k), float(alpha), ptr(r1), ptr(r2), ptr(gold_col), ptr(gold_val),
                                    ptr(cnt), ptr(diag), ptr(topv), ptr(topi), ptr(ws), ws_bytes, stream_ptr()), "jmac_sim_topk")
    return {"topv": topv, "topi": topi, "count": cnt, "diag": diag}


def csls_gold(alpha, diag, r1, r2, gold_col):
    lib = _lib.load()
    out = torch.empty_like(diag)
    with torch.cuda.device(diag.device):
        check(lib.jmac_csls_gold(float(alpha), ptr(diag), ptr(r1), ptr(r2), ptr(gold_col), int(diag.numel()), ptr(out), stream_ptr()), "jmac_csls_gold")
    return out


def _ranks_from_tables(A: SimTable, B: SimTable, csls_k: int, gold_col: torch.Tensor):
    """Passes A, B, C on prepared tables; gold column of row i = gold_col[i].  Returns (ranks, top1)."""
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


def alignment_ranks(embed1, embed2, metric="inner", normalize=False, csls_k=0, device=None, precision: Optional[str] = None):
    """0-based rank of the gold column (row i <-> column i) for every row, plus the top-1 column.
    The streaming equivalent of sim() + csls_sim() + calculate_rank(accurate=True)."""
    e1, e2 = _dev_f32(embed1, device), _dev_f32(embed2, device)
    if metric not in ("inner", "cosine"):
        raise NotImplementedError(f"metric {metric!r} has no streaming form; greedy_alignment() serves it through the dense path")
    norm = bool(normalize) or metric == "cosine"
    A, B = to_rows(e1, normalize=norm, precision=precision), to_rows(e2, normalize=norm, precision=precision)
    gold_col = torch.arange(A.shape[0], dtype=torch.int32, device=A.device)
    return _ranks_from_tables(A, B, csls_k, gold_col)


def alignment_ranks_bf16(A, B, csls_k=10, group=None):
    """Same three passes on already-converted tables ([n, dpad] bf16 tensors or SimTables, gold column of row i = i).
    With ``group`` the work is ROW-SHARDED by source entity across the ranks (every rank holds both
    tables, 512 MB each at 1M x 256): pass A and C on this rank's rows of A, pass B on this rank's rows
    of B, then an all-gather of r2 and a final all-gather of the per-row results (rows are disjoint,
    so the "top-k merge" is a concatenation).  Returns (ranks[n1], top1[n1]) on every rank."""
    import torch.distributed as dist
    A, B = _table(A), _table(B)
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


def _dense_ranks(s: torch.Tensor):
    """Rank of the gold column by counting on a dense similarity matrix (precomputed ``simi_matrix`` or a metric
    without a streaming form)."""
    n = s.shape[0]
    gold = s[torch.arange(n, device=s.device), torch.arange(n, device=s.device)]
    return (s > gold[:, None]).sum(1).to(torch.int32), s.argmax(1).to(torch.int32)


def greedy_alignment(embed1, embed2, top_k, nums_threads, metric, normalize, csls_k, accurate, simi_matrix=None, return_simi=False,
                     logger=None, precision: Optional[str] = None):
    """modules/finding/alignment.py:10-69 -- returns (top_k, hits% rounded to 3 dp, mr, mrr).
    ``accurate=False`` (argpartition, alignment.py:101) only guarantees the order of the top-k positions in the
    reference; Hits@k are identical, MR / MRR are computed from the exact ranks here."""
    assert 1 in top_k                                              # alignment.py:91
    if simi_matrix is not None:
        ranks, _ = _dense_ranks(_dev_f32(simi_matrix))             # alignment.py:40-41: the matrix replaces sim()
    elif metric in ("inner", "cosine"):
        ranks, _ = alignment_ranks(embed1, embed2, metric=metric, normalize=normalize, csls_k=csls_k, precision=precision)
    else:
        ranks, _ = _dense_ranks(torch.from_numpy(sim(embed1, embed2, metric, normalize, csls_k)).to(_DEF_DEV))
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


def get_neg(ILL, emb_src, emb_dst, k, precision: Optional[str] = None):
    """modules/utils/util.py:31-54 -- indices of the k most similar rows of emb_dst for each seed,
    flattened to [len(ILL) * k] (LongTensor, on the CPU like the reference's inputs when they are)."""
    src, dst = _dev_f32(emb_src), _dev_f32(emb_dst)
    A = to_rows(src, idx=ILL, precision=precision)
    B = to_rows(dst, precision=precision)
    res = sim_topk(A, B, int(k))
    out = res["topi"].to(torch.int64).reshape(-1)
    if torch.is_tensor(emb_src) and not emb_src.is_cuda:
        out = out.cpu()
    return out


# ---- alignment quality / seed enlargement (train.py:231-259, 163-169) ---------------------------
def softmax_stats(A, B, alpha, row_mask=None, col_mask=None, masked_value=None, split: bool = True):
    """Per-row softmax statistics of alpha * A.B^T (see jmac_sim_softmax_stats): dict(max, arg, Z, W)."""
    lib = _lib.load()
    A, B = _table(A), _table(B)
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
    ws, ws_bytes = _workspace(lib, n1, n2, 0, dev) if split else (None, 0)
    mv = float(-alpha if masked_value is None else masked_value)
    with torch.cuda.device(dev):
        if A.f32:
            check(lib.jmac_sim_softmax_stats_f32(ptr(A.hi), ptr(A.lo), ptr(B.hi), ptr(B.lo), n1, n2, dpad, float(alpha), ptr(row_mask),
                                                 ptr(col_mask), mv, ptr(out["max"]), ptr(out["arg"]), ptr(out["Z"]), ptr(out["W"]), ptr(ws),
                                                 ws_bytes, stream_ptr()), "jmac_sim_softmax_stats_f32")
        else:
            check(lib.jmac_sim_softmax_stats(ptr(A.hi), ptr(B.hi), n1, n2, dpad, float(alpha), ptr(row_mask), ptr(col_mask), mv,
                                             ptr(out["max"]), ptr(out["arg"]), ptr(out["Z"]), ptr(out["W"]), ptr(ws), ws_bytes, stream_ptr()),
                  "jmac_sim_softmax_stats")
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


def compute_alignment_quality(embedding1, embedding2, list1, list2, scale=20.0, precision: Optional[str] = None):
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
    A_sub, B_sub = to_rows(e1, idx=l1, precision=precision), to_rows(e2, idx=l2, precision=precision)
    s12 = softmax_stats(A_sub, B_sub, scale)
    s21 = softmax_stats(B_sub, A_sub, scale)
    ent = (torch.log(s12["Z"]) - s12["W"] / s12["Z"]).mean() + (torch.log(s21["Z"]) - s21["W"] / s21["Z"]).mean()
    # (2) full matrix with rows not in list1 / columns not in list2 set to -1 (train.py:252-255)
    A, B = to_rows(e1, precision=precision), to_rows(e2, precision=precision)
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
    """modules/finding/similarity.py:13-55 for callers that really want the dense matrix (numpy float32 [n1,n2]).
    Library ops on the device (fp32 matmul; float64 for the cdist-based metrics as in the reference); the
    evaluation path uses alignment_ranks() instead."""
    e1, e2 = _dev_f32(embed1), _dev_f32(embed2)
    if normalize:                                                   # sklearn.preprocessing.normalize (:34-36)
        e1 = torch.nn.functional.normalize(e1, 2, -1)
        e2 = torch.nn.functional.normalize(e2, 2, -1)
    if metric == "inner" or (metric == "cosine" and normalize):
        s = e1 @ e2.t()
    elif metric == "cosine":                                        # 1 - cdist(cosine) in float64 (:44-46)
        a, b = e1.double(), e2.double()
        s = (1.0 - (1.0 - (a @ b.t()) / (a.norm(dim=1)[:, None] * b.norm(dim=1)[None, :]))).float()
    elif metric == "euclidean":                                     # 1 - euclidean_distances (:41-43)
        s = (1.0 - torch.cdist(e1.double(), e2.double(), p=2)).float()
    elif metric in ("manhattan", "cityblock"):                      # 1 - cdist(cityblock) (:47-49)
        s = (1.0 - torch.cdist(e1.double(), e2.double(), p=1)).float()
    else:
        raise NotImplementedError(f"metric {metric!r}")
    if csls_k > 0:
        s = _csls_dense(s, csls_k)
    return s.cpu().numpy()


def _csls_dense(s: torch.Tensor, k: int) -> torch.Tensor:
    r1 = s.topk(k, dim=1)[0].mean(1)
    r2 = s.topk(k, dim=0)[0].mean(0)
    return (2 * s.t() - r1).t() - r2                                # operation order of similarity.py:74-77


def csls_sim(sim_mat, k):
    """modules/finding/similarity.py:58-78 on a dense matrix."""
    return _csls_dense(_dev_f32(sim_mat), k).cpu().numpy()
