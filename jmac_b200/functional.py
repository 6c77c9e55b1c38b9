"""Host-side glue between PyTorch tensors and the C ABI: the graph-plan cache and the
``torch.autograd.Function`` wrappers around the sm_100a kernels.  PyTorch is plumbing here (device
memory, streams, autograd bookkeeping, the node-level library GEMMs); every per-edge / per-triple
operation runs in libjmac_b200.so.  No CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.JmacError("jmac_b200 kernels need CUDA tensors; there is no CPU fallback")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def as_index(x, device) -> torch.Tensor:
    """Normalise the index types the reference passes around (python lists, numpy int64, numpy
    float64 such as ``1951.0`` from train.py:189-192, LongTensors) to a contiguous int64 device
    tensor, once, at the boundary (SURVEY 8c 'input-type quirks')."""
    if torch.is_tensor(x):
        t = x
    else:
        a = np.asarray(x)
        if a.dtype.kind == "f":
            a = a.astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(a.astype(np.int64, copy=False)))
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t.to(device, non_blocking=True).contiguous()


# ----------------------------------------------------------------------------------------------
# graph plan
# ----------------------------------------------------------------------------------------------
class GraphPlan:
    """Owns a ``jmac_plan`` (CSR-by-destination + by-source + by-relation orderings)."""

    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, num_ent: int, n_rel: int, chunk: int = 0, n_src: int = None):
        _need_cuda(edge_index, edge_type)
        lib = _lib.load()
        ei = edge_index.to(torch.int64).contiguous()
        et = edge_type.to(torch.int64).contiguous()
        if ei.dim() != 2 or ei.shape[0] != 2 or et.shape[0] != ei.shape[1]:
            raise ValueError("edge_index must be [2,E] and edge_type [E]")
        self.E, self.N, self.n_rel = int(ei.shape[1]), int(num_ent), int(n_rel)
        self.N_src = int(n_src) if n_src is not None else self.N      # != N only for destination-sharded plans
        self.device = ei.device
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.jmac_plan_create_bipartite(ptr(ei[0]), ptr(ei[1]), ptr(et), self.E, self.N, self.N_src, self.n_rel, int(chunk),
                                                 stream_ptr(), ctypes.byref(handle)), "jmac_plan_create")
        self._h = handle
        info = (ctypes.c_int64 * 8)()
        check(lib.jmac_plan_info(self._h, info), "jmac_plan_info")
        self.info = list(info)
        self.max_in_degree = int(info[7])

    @property
    def handle(self):
        return self._h

    def export(self):
        """(rowptr, csr_src, csr_type, csr_eid, sqrt_deg) as device tensors -- for tests."""
        lib = _lib.load()
        rowptr = torch.empty(self.N + 1, dtype=torch.int32, device=self.device)
        src = torch.empty(self.E, dtype=torch.int32, device=self.device)
        typ = torch.empty(self.E, dtype=torch.int32, device=self.device)
        eid = torch.empty(self.E, dtype=torch.int32, device=self.device)
        sdeg = torch.empty(self.N, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.jmac_plan_export(self._h, ptr(rowptr), ptr(src), ptr(typ), ptr(eid), ptr(sdeg), stream_ptr()), "jmac_plan_export")
        return rowptr, src, typ, eid, sdeg

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().jmac_plan_destroy(h)
            except Exception:
                pass


_PLAN_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_PLAN_CACHE_SIZE = 16
_PLAN_PINS: Optional[list] = None      # while a CUDA graph is being captured (jmac_b200.graphs): every plan handed out is also kept here,
                                       # so that the captured kernels' plan arrays outlive an eviction from the LRU cache


def get_plan(edge_index: torch.Tensor, edge_type: torch.Tensor, num_ent: int, n_rel: int) -> GraphPlan:
    """Plan cache keyed on tensor identity + version: the reference reuses one edge set for
    thousands of optimiser steps (train.py:447-485) and rebuilds tensors when the graph changes."""
    key = _plan_key(edge_index, edge_type, num_ent, n_rel)
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        _PLAN_CACHE.move_to_end(key)
        if _PLAN_PINS is not None:
            _PLAN_PINS.append(hit[0])
        return hit[0]
    plan = GraphPlan(edge_index, edge_type, num_ent, n_rel)
    # keep the key tensors alive so that their addresses cannot be recycled while cached
    _PLAN_CACHE[key] = (plan, edge_index, edge_type)
    while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
        _PLAN_CACHE.popitem(last=False)
    if _PLAN_PINS is not None:
        _PLAN_PINS.append(plan)
    return plan


def _plan_key(edge_index, edge_type, num_ent, n_rel):
    return (edge_index.data_ptr(), edge_type.data_ptr(), tuple(edge_index.shape), edge_index._version, edge_type._version,
            int(num_ent), int(n_rel), str(edge_index.device))


def register_plan(edge_index: torch.Tensor, edge_type: torch.Tensor, num_ent: int, n_rel: int, plan: GraphPlan):
    """Put a plan that was built ahead of time (e.g. on a copy stream while the previous step computes) into the cache under
    the key the layers will look it up with."""
    _PLAN_CACHE[_plan_key(edge_index, edge_type, num_ent, n_rel)] = (plan, edge_index, edge_type)
    while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
        _PLAN_CACHE.popitem(last=False)


def get_edge_table_plan(edge_index: torch.Tensor, num_ent: int) -> GraphPlan:
    """Plan for per-edge source tables (comp_op='mult'): destination = edge_index[0], "source" of edge e = row e of an
    [E,d] table, a single (zero) relation.  Cached like get_plan."""
    key = ("edge-table", edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_ent), str(edge_index.device))
    hit = _PLAN_CACHE.get(key)
    if hit is not None:
        _PLAN_CACHE.move_to_end(key)
        return hit[0]
    dev = edge_index.device
    E = int(edge_index.shape[1])
    ei = torch.stack([edge_index[0].to(torch.int64), torch.arange(E, dtype=torch.int64, device=dev)])
    plan = GraphPlan(ei, torch.zeros(E, dtype=torch.int64, device=dev), num_ent, 1, n_src=max(E, 1))
    _PLAN_CACHE[key] = (plan, edge_index)
    while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
        _PLAN_CACHE.popitem(last=False)
    return plan


def clear_plan_cache():
    _PLAN_CACHE.clear()


# ----------------------------------------------------------------------------------------------
# edge phase
# ----------------------------------------------------------------------------------------------
class EdgePhaseFn(torch.autograd.Function):
    """agg[i] = sqrt(deg_i) * sum_e softmax_i(a . lrelu(P[i]+Q[j]-S[t])) * (X[j] - Rp[t]).
    With ``Xself`` ([N_dst,d], the destinations' own rows) the self-loop propagate is fused and the
    result is T = agg + Xself - Rp[-1] (src/jmac_model.py:49-52 before the message GEMM)."""

    @staticmethod
    def forward(ctx, plan: GraphPlan, P, Q, X, S, Rp, a_att, slope: float, Xself=None):
        _need_cuda(P, Q, X, S, Rp, a_att)
        lib = _lib.load()
        P, Q, X, S, Rp, a = (_f32c(t) for t in (P, Q, X, S, Rp, a_att.reshape(-1)))
        N, d = P.shape
        if N != plan.N or X.shape[0] != plan.N_src or Q.shape[0] != plan.N_src or S.shape[0] != plan.n_rel or Rp.shape[0] != plan.n_rel:
            raise ValueError("table sizes do not match the graph plan")
        agg = torch.empty_like(P)
        lse = torch.empty(N, dtype=torch.float32, device=X.device)
        T = None
        if Xself is not None:
            Xself = _f32c(Xself)
            if Xself.shape != P.shape:
                raise ValueError("Xself must be [N_dst, d]")
            T = torch.empty_like(P)
        # per-edge record (logit + sign mask, later {w, ds}) only when a gradient will be asked for
        need_grad = any(ctx.needs_input_grad)
        erec = None
        if need_grad:
            erec = torch.empty(max(plan.E, 1) * int(lib.jmac_edge_record_bytes(d)), dtype=torch.uint8, device=X.device)
        with torch.cuda.device(X.device):
            check(lib.jmac_edge_fwd_fused(plan.handle, ptr(P), ptr(Q), ptr(X), ptr(S), ptr(Rp), ptr(a), float(slope), d, ptr(agg), ptr(lse),
                                          ptr(Xself), ptr(T), ptr(erec), stream_ptr()), "jmac_edge_fwd")
        ctx.plan, ctx.slope, ctx.fused = plan, float(slope), Xself is not None
        if need_grad:
            ctx.save_for_backward(P, Q, X, S, Rp, a, agg, lse, erec)
        ctx.a_shape = a_att.shape
        return T if T is not None else agg

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        P, Q, X, S, Rp, a, agg, lse, erec = ctx.saved_tensors
        plan, slope = ctx.plan, ctx.slope
        d = X.shape[1]
        G = _f32c(g)
        dev = X.device
        dP = torch.empty_like(P)
        dXe = torch.empty_like(X)
        dQ = torch.empty_like(X)
        dRp = torch.empty_like(Rp)
        dS = torch.empty_like(S)
        acc = torch.zeros(2, d, dtype=torch.float32, device=dev)      # [0] = d a_att, [1] = column sums of G (fused self loop)
        da, gsum = acc[0], (acc[1] if ctx.fused else None)
        with torch.cuda.device(dev):
            st = stream_ptr()
            check(lib.jmac_edge_bwd_dst(plan.handle, ptr(P), ptr(X), ptr(Rp), ptr(a), slope, d, ptr(agg), ptr(lse), ptr(G), ptr(dP),
                                        ptr(erec), ptr(da), ptr(gsum), st), "jmac_edge_bwd_dst")
            check(lib.jmac_edge_bwd_src(plan.handle, ptr(G), ptr(erec), ptr(a), slope, d, ptr(Q), ptr(dXe), ptr(dQ), ptr(da), st),
                  "jmac_edge_bwd_src")
            check(lib.jmac_edge_bwd_rel(plan.handle, ptr(G), ptr(erec), ptr(a), slope, d, ptr(S), ptr(dRp), ptr(dS), ptr(da), st),
                  "jmac_edge_bwd_rel")
        dXself = None
        if ctx.fused:
            dXself = G                                  # T = agg + Xself - Rp[-1]
            dRp[-1] -= gsum
        return None, dP, dQ, dXe, dS, dRp, da.view(ctx.a_shape), None, dXself


def edge_phase(plan, P, Q, X, S, Rp, a_att, slope, Xself=None):
    return EdgePhaseFn.apply(plan, P, Q, X, S, Rp, a_att, slope, Xself)


class AttnEdgeFn(torch.autograd.Function):
    """The layer-input part of RelationAwareLayer.forward in one autograd node (comp_op='sub'):
    ``P = Xd W_top, Q = Xs W_bot, S = R' W_bot`` and the edge phase with the fused self loop, returning
    ``T = agg + Xd - R'[loop]``.  ``Xd`` are the destination rows, ``Xs`` the source table (``None``: the same table,
    single GPU; destination-sharded: the all-gathered table).  Same kernels as ``node_mm`` + ``edge_phase``; what the
    fusion buys is the backward: ``dXd = dP W_top^T + dT`` and ``dXs = dQ W_bot^T + dXe`` (one sum when the tables
    coincide) are accumulated in the epilogues of the input-gradient GEMMs instead of elementwise passes by autograd."""

    @staticmethod
    def forward(ctx, plan: GraphPlan, Xd, Xs, W_att, Rp, a_att, slope: float):
        same = Xs is None
        _need_cuda(Xd, W_att, Rp, a_att)
        lib = _lib.load()
        Xd, W_att, Rp, a = _f32c(Xd), _f32c(W_att), _f32c(Rp), _f32c(a_att.reshape(-1))
        Xs = Xd if same else _f32c(Xs)
        N, d = Xd.shape
        if N != plan.N or Xs.shape[0] != plan.N_src or Rp.shape[0] != plan.n_rel:
            raise ValueError("table sizes do not match the graph plan")
        Wt, Wb = W_att[:d], W_att[d:]
        P = gemm_nn(Xd, Wt)
        Q = gemm_nn(Xs, Wb)
        S = small_mm(Rp, Wb)
        agg = torch.empty_like(Xd)
        T = torch.empty_like(Xd)
        lse = torch.empty(N, dtype=torch.float32, device=Xd.device)
        need_grad = any(ctx.needs_input_grad)
        erec = torch.empty(max(plan.E, 1) * int(lib.jmac_edge_record_bytes(d)), dtype=torch.uint8, device=Xd.device) if need_grad else None
        with torch.cuda.device(Xd.device):
            check(lib.jmac_edge_fwd_fused(plan.handle, ptr(P), ptr(Q), ptr(Xs), ptr(S), ptr(Rp), ptr(a), float(slope), d, ptr(agg), ptr(lse),
                                          ptr(Xd), ptr(T), ptr(erec), stream_ptr()), "jmac_edge_fwd")
        ctx.plan, ctx.slope, ctx.a_shape, ctx.same = plan, float(slope), a_att.shape, same
        if need_grad:
            ctx.save_for_backward(Xd, Xs, W_att, Rp, a, P, Q, S, agg, lse, erec)
        return T

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        Xd, Xs, W_att, Rp, a, P, Q, S, agg, lse, erec = ctx.saved_tensors
        plan, slope, same = ctx.plan, ctx.slope, ctx.same
        d = Xd.shape[1]
        G = _f32c(g)
        dev = Xd.device
        Wt, Wb = W_att[:d], W_att[d:]
        dP, dXe, dQ = torch.empty_like(Xd), torch.empty_like(Xs), torch.empty_like(Xs)
        dRp, dS = torch.empty_like(Rp), torch.empty_like(S)
        acc = torch.zeros(2, d, dtype=torch.float32, device=dev)      # [0] = d a_att, [1] = column sums of G
        da, gsum = acc[0], acc[1]
        with torch.cuda.device(dev):
            st = stream_ptr()
            check(lib.jmac_edge_bwd_dst(plan.handle, ptr(P), ptr(Xs), ptr(Rp), ptr(a), slope, d, ptr(agg), ptr(lse), ptr(G), ptr(dP),
                                        ptr(erec), ptr(da), ptr(gsum), st), "jmac_edge_bwd_dst")
            check(lib.jmac_edge_bwd_src(plan.handle, ptr(G), ptr(erec), ptr(a), slope, d, ptr(Q), ptr(dXe), ptr(dQ), ptr(da), st),
                  "jmac_edge_bwd_src")
            check(lib.jmac_edge_bwd_rel(plan.handle, ptr(G), ptr(erec), ptr(a), slope, d, ptr(S), ptr(dRp), ptr(dS), ptr(da), st),
                  "jmac_edge_bwd_rel")
        dRp[-1] -= gsum                                       # T = agg + Xd - Rp[-1]: column sums of G, picked up by the destination pass
        small_mm_nt_acc(dS, Wb, dRp)                          # S = Rp Wb: dRp += dS Wb^T
        dXd = dXs = dW = None
        if same:
            if ctx.needs_input_grad[1]:
                dXd = gemm_nt(dP, Wt, add1=dXe, add2=G, out=dXe)                  # dP Wt^T + dXe + dT, in place over dXe
                dXd = gemm_nt(dQ, Wb, add1=dXd, out=dXd)                          # + dQ Wb^T
        else:
            if ctx.needs_input_grad[1]:
                dXd = gemm_nt(dP, Wt, add1=G)                                     # dP Wt^T + dT
            if ctx.needs_input_grad[2]:
                dXs = gemm_nt(dQ, Wb, add1=dXe, out=dXe)                          # dQ Wb^T + dXe: the partial that is reduce-scattered
        if ctx.needs_input_grad[3]:
            dW = torch.empty_like(W_att)
            dW[:d] = gemm_tn(Xd, dP)
            dWb = gemm_tn(Xs, dQ)
            dWb += small_mm_tn(Rp, dS)
            dW[d:] = dWb
        return None, dXd, dXs, dW, dRp, da.view(ctx.a_shape), None


def attn_edge_eligible(X: torch.Tensor, d: int) -> bool:
    return X.is_cuda and X.dtype == torch.float32 and _tc_eligible(X.shape[0], d, d) and _tn_eligible(X.shape[0], d, d)


def attn_edge(plan, X, W_att, Rp, a_att, slope, X_src=None):
    """T = agg + X - R'[loop] for destination rows ``X`` (sources: ``X_src``, default the same table)."""
    return AttnEdgeFn.apply(plan, X, X_src, W_att, Rp, a_att, slope)


# ----------------------------------------------------------------------------------------------
# node-level dense products (3xTF32 on tcgen05, fp32-accurate)
# ----------------------------------------------------------------------------------------------
def _split_weight(W: torch.Tensor, transpose: bool):
    """The small operand of a node product, pre-split for the 3xTF32 kernel by ONE launch: hi = W rounded to nearest TF32
    (ties away, cvt.rna.tf32.f32), lo = W - hi (exact in fp32), laid out as W or as W^T (contiguous)."""
    lib = _lib.load()
    if W.dtype != torch.float32:
        raise TypeError(f"expected float32, got {W.dtype}")
    if W.stride(1) != 1:
        W = W.contiguous()
    rows, cols = W.shape
    shape = (cols, rows) if transpose else (rows, cols)
    hi = torch.empty(shape, dtype=torch.float32, device=W.device)
    lo = torch.empty(shape, dtype=torch.float32, device=W.device)
    with torch.cuda.device(W.device):
        check(lib.jmac_split_tf32(ptr(W), rows, cols, W.stride(0), int(bool(transpose)), ptr(hi), ptr(lo), stream_ptr()), "jmac_split_tf32")
    return hi, lo


def _gemm_presplit(A: torch.Tensor, hi: torch.Tensor, lo: torch.Tensor, add1=None, add2=None, out=None) -> torch.Tensor:
    lib = _lib.load()
    A = _f32c(A)
    M, K = A.shape
    N = hi.shape[0]
    C = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=A.device)
    for t in (add1, add2, out):
        if t is not None and (t.shape != (M, N) or t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("addends / out must be contiguous float32 [M,N]")
    if add1 is None and add2 is not None:
        add1, add2 = add2, None
    with torch.cuda.device(A.device):
        check(lib.jmac_gemm_tf32x3_add(ptr(A), K, ptr(hi), ptr(lo), K, ptr(C), N, M, N, K, ptr(add1), ptr(add2), N, stream_ptr()),
              "jmac_gemm_tf32x3_add")
    return C


def gemm_nt(A: torch.Tensor, B_nk: torch.Tensor, add1=None, add2=None, out=None) -> torch.Tensor:
    """C[M,N] = A[M,K] @ B_nk[N,K]^T (+ add1 + add2) through jmac_gemm_tf32x3_add (raises on unsupported shapes).
    ``add1``/``add2``: [M,N] contiguous addends folded into the epilogue; ``out`` may be one of them (in-place accumulate)."""
    _need_cuda(A, B_nk)
    if B_nk.shape[1] != A.shape[1]:
        raise ValueError("inner dimensions differ")
    hi, lo = _split_weight(B_nk, False)
    return _gemm_presplit(A, hi, lo, add1, add2, out)


def gemm_nn(A: torch.Tensor, W_kn: torch.Tensor, add1=None, add2=None, out=None) -> torch.Tensor:
    """C[M,N] = A[M,K] @ W_kn[K,N] (+ addends): the transposition of the weight is folded into its split pass."""
    _need_cuda(A, W_kn)
    if W_kn.shape[0] != A.shape[1]:
        raise ValueError("inner dimensions differ")
    hi, lo = _split_weight(W_kn, True)
    return _gemm_presplit(A, hi, lo, add1, add2, out)


def gemm_tn(X: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    """D[KO,N] = X[M,KO]^T @ G[M,N] (weight gradients: reduction over the node rows) through
    jmac_gemm_tn_tf32x3."""
    lib = _lib.load()
    _need_cuda(X, G)
    X, G = _f32c(X), _f32c(G)
    M, KO = X.shape
    N = G.shape[1]
    D = torch.empty(KO, N, dtype=torch.float32, device=X.device)
    if KO <= 256 and N <= 256:
        ws = torch.empty(int(lib.jmac_gemm_tn_workspace_bytes(KO, N)) // 4, dtype=torch.float32, device=X.device)
        with torch.cuda.device(X.device):
            check(lib.jmac_gemm_tn_tf32x3(ptr(X), KO, ptr(G), N, ptr(D), M, KO, N, ptr(ws), stream_ptr()), "jmac_gemm_tn_tf32x3")
        return D
    # wider tables (JMAC_DBPv1's d = 300 runs at the padded width 320): one launch per block of <= 256 columns of X and of G
    # (strided views of the row-major tables), each writing its own [KOb, Nb] block
    ws = torch.empty(int(lib.jmac_gemm_tn_workspace_bytes(256, 256)) // 4, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        for k0 in range(0, KO, 256):
            kb = min(256, KO - k0)
            for n0 in range(0, N, 256):
                nb = min(256, N - n0)
                blk = torch.empty(kb, nb, dtype=torch.float32, device=X.device)
                check(lib.jmac_gemm_tn_tf32x3(X.data_ptr() + 4 * k0, KO, G.data_ptr() + 4 * n0, N, ptr(blk), M, kb, nb, ptr(ws), stream_ptr()),
                      "jmac_gemm_tn_tf32x3")
                D[k0:k0 + kb, n0:n0 + nb] = blk
    return D


# Measured on the DBP-5L-shaped step (profiles/c2_step_launches_r02.md): for the 962-row relation tables the tcgen05 kernel (pipeline
# fill, TMEM allocation, a separate split launch) took 18 us against 15 us for the library's SIMT product, and its weight-gradient
# twin 31 us against 10 us -- node tables (>= 4096 rows) go to the tensor cores, relation tables stay with the library.
_TC_MIN_ROWS = 4096


def small_mm(A: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """A @ W for the relation tables (a few hundred to a few thousand rows): the tensor-core kernel when its tile grid pays,
    the library product otherwise.  No autograd (used inside the autograd nodes)."""
    if _tc_eligible(A.shape[0], A.shape[1], W.shape[1]):
        return gemm_nn(A, W)
    return A @ W


def small_mm_nt_acc(A: torch.Tensor, W: torch.Tensor, acc: torch.Tensor) -> torch.Tensor:
    """acc += A @ W^T, in place."""
    if _tc_eligible(A.shape[0], A.shape[1], W.shape[0]) and acc.is_contiguous():
        return gemm_nt(A, W, add1=acc, out=acc)
    acc += A @ W.t()
    return acc


def small_mm_tn(X: torch.Tensor, G: torch.Tensor) -> torch.Tensor:
    """X^T @ G."""
    if _tn_eligible(X.shape[0], X.shape[1], G.shape[1]):
        return gemm_tn(X, G)
    return X.t() @ G


def _tn_eligible(M, KO, N) -> bool:
    return M >= _TC_MIN_ROWS and KO % 32 == 0 and 32 <= KO <= 512 and N % 32 == 0 and 32 <= N <= 512


def _tc_eligible(M, K, N) -> bool:
    return M >= _TC_MIN_ROWS and N % 32 == 0 and 32 <= N <= 512 and K % 16 == 0 and K >= 16


def padded_width(d: int) -> int:
    """Feature width the tensor-core node products and the edge kernels run at for a layer of width ``d``: the next multiple of
    32 (JMAC_DBPv1's d = 300 -> 320); zero columns stay exactly zero through the layer (zero weight rows / columns, zero BN affine)."""
    return (int(d) + 31) // 32 * 32


class NodeMatmulFn(torch.autograd.Function):
    """Y = X @ W for a tall node table X[M,K] and a small weight W[K,N]."""

    @staticmethod
    def forward(ctx, X, W):
        ctx.save_for_backward(X, W)
        return gemm_nn(X, W)

    @staticmethod
    def backward(ctx, G):
        X, W = ctx.saved_tensors
        G = _f32c(G)
        dX = dW = None
        if ctx.needs_input_grad[0]:
            M, N = G.shape
            dX = gemm_nt(G, W) if _tc_eligible(M, N, W.shape[0]) else G @ W.t()
        if ctx.needs_input_grad[1]:
            dW = gemm_tn(X, G) if _tn_eligible(X.shape[0], X.shape[1], G.shape[1]) else X.t() @ G
        return dX, dW


def node_mm(X: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    """``torch.mm(X, W)`` semantics; large node tables go through the 3xTF32 tensor-core kernel."""
    if X.is_cuda and X.dtype == torch.float32 and X.dim() == 2 and _tc_eligible(X.shape[0], X.shape[1], W.shape[1]):
        return NodeMatmulFn.apply(X, W)
    return torch.mm(X, W)


class ConcatLinearFn(torch.autograd.Function):
    """``torch.mm(torch.cat(xs, 1), W)`` without the concatenation: sum_i xs[i] @ W[rows of block i], every product after the
    first accumulated in the GEMM epilogue of the next (src/jmac_model.py:185,194,203: the name / alignment channel)."""

    @staticmethod
    def forward(ctx, W, *xs):
        xs = [_f32c(x) for x in xs]
        W = _f32c(W)
        out, off = None, 0
        for x in xs:
            k = x.shape[1]
            out = gemm_nn(x, W[off:off + k]) if out is None else gemm_nn(x, W[off:off + k], add1=out, out=out)
            off += k
        ctx.save_for_backward(W, *xs)
        return out

    @staticmethod
    def backward(ctx, G):
        W, *xs = ctx.saved_tensors
        G = _f32c(G)
        dW = torch.empty_like(W) if ctx.needs_input_grad[0] else None
        dxs, off = [], 0
        for i, x in enumerate(xs):
            k = x.shape[1]
            dxs.append(gemm_nt(G, W[off:off + k]) if ctx.needs_input_grad[1 + i] else None)
            if dW is not None:
                dW[off:off + k] = gemm_tn(x, G)
            off += k
        return (dW, *dxs)


def concat_linear(xs: Sequence[torch.Tensor], W: torch.Tensor) -> torch.Tensor:
    """``torch.mm(torch.cat(xs, dim=1), W)``; large node tables skip the [N, sum k_i] concatenation."""
    M, N = xs[0].shape[0], W.shape[1]
    ok = all(x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and _tc_eligible(M, x.shape[1], N) and _tn_eligible(M, x.shape[1], N)
             and _tc_eligible(M, N, x.shape[1]) for x in xs)
    if ok and sum(x.shape[1] for x in xs) == W.shape[0]:
        return ConcatLinearFn.apply(W, *xs)
    return torch.mm(torch.cat(list(xs), dim=1), W)


class RowNormalizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        lib = _lib.load()
        x = _f32c(x)
        N, d = x.shape
        y = torch.empty_like(x)
        need = ctx.needs_input_grad[0]
        inv = torch.empty(N, dtype=torch.float32, device=x.device) if need else None
        with torch.cuda.device(x.device):
            check(lib.jmac_row_normalize_fwd(ptr(x), N, d, float(eps), ptr(y), ptr(inv), stream_ptr()), "jmac_row_normalize_fwd")
        if need:
            ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        y, inv = ctx.saved_tensors
        g = _f32c(g)
        dx = torch.empty_like(y)
        with torch.cuda.device(y.device):
            check(lib.jmac_row_normalize_bwd(ptr(y), ptr(inv), ptr(g), y.shape[0], y.shape[1], ptr(dx), stream_ptr()), "jmac_row_normalize_bwd")
        return dx, None


def row_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """``F.normalize(x, 2, -1)`` for a [N, d] table (one kernel each way); other shapes / dtypes go to the library op."""
    if x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] % 4 == 0 and 4 <= x.shape[1] <= 512:
        return RowNormalizeFn.apply(x, eps)
    return torch.nn.functional.normalize(x, 2, -1, eps)


# ----------------------------------------------------------------------------------------------
# BatchNorm1d(over nodes) + tanh
# ----------------------------------------------------------------------------------------------
class BNTanhFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pre, gamma, beta, running_mean, running_var, momentum, eps, training, group, n_global=None):
        _need_cuda(pre, gamma, beta)
        lib = _lib.load()
        pre, gamma, beta = _f32c(pre), _f32c(gamma), _f32c(beta)
        N, d = pre.shape
        dev = pre.device
        out = torch.empty_like(pre)
        mean = torch.empty(d, dtype=torch.float32, device=dev)
        rstd = torch.empty(d, dtype=torch.float32, device=dev)
        n_total = N
        sums = None
        with torch.cuda.device(dev):
            st = stream_ptr()
            if training:
                sums = torch.zeros(2 * d + 1, dtype=torch.float64, device=dev)
                check(lib.jmac_bn_stats(ptr(pre), N, d, ptr(sums), st), "jmac_bn_stats")
                if group is not None:  # BN is over ALL nodes: all-reduce (sum x, sum x^2) across destination shards
                    if n_global is None:
                        raise ValueError("sharded BatchNorm needs the global row count (a constant of the partition: Shard.n_total)")
                    torch.distributed.all_reduce(sums, group=group)
                    n_total = int(n_global)
                st = stream_ptr()
            check(lib.jmac_bn_tanh_apply(ptr(pre), N, d, ptr(sums), n_total, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var),
                                         float(momentum), float(eps), int(bool(training)), ptr(out), ptr(mean), ptr(rstd), st),
                  "jmac_bn_tanh_apply")
        ctx.save_for_backward(pre, out, gamma, mean, rstd)
        ctx.training, ctx.n_total, ctx.group = bool(training), n_total, group
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        pre, out, gamma, mean, rstd = ctx.saved_tensors
        N, d = pre.shape
        dev = pre.device
        gout = _f32c(gout)
        bsums = torch.zeros(2 * d, dtype=torch.float64, device=dev)
        gpre = torch.empty_like(pre)
        dgamma = torch.empty(d, dtype=torch.float32, device=dev)
        dbeta = torch.empty(d, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = stream_ptr()
            check(lib.jmac_bn_tanh_bwd_stats(ptr(gout), ptr(out), ptr(pre), N, d, ptr(mean), ptr(rstd), ptr(bsums), st), "jmac_bn_tanh_bwd_stats")
            if ctx.group is not None and ctx.training:
                # the per-shard (sum dy, sum dy*xhat) are also the shard's contribution to dbeta/dgamma;
                # the statistics coupling needs the global sums, the parameter grads are all-reduced by the caller
                local = bsums.clone()
                torch.distributed.all_reduce(bsums, group=ctx.group)
            else:
                local = bsums
            check(lib.jmac_bn_tanh_bwd_apply(ptr(gout), ptr(out), ptr(pre), N, d, ptr(mean), ptr(rstd), ptr(gamma), ptr(bsums),
                                             ctx.n_total, int(ctx.training), ptr(gpre), None, None, st), "jmac_bn_tanh_bwd_apply")
        dbeta = local[:d].to(torch.float32)
        dgamma = local[d:].to(torch.float32)
        return gpre, dgamma, dbeta, None, None, None, None, None, None, None


def bn_tanh(pre, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5, training=True, group=None, n_global=None):
    """tanh(BatchNorm1d(pre)) over the node dimension.  ``group`` + ``n_global``: the rows are one destination shard of
    ``n_global`` nodes; the per-feature sums are all-reduced over ``group`` (no host synchronisation)."""
    return BNTanhFn.apply(pre, gamma, beta, running_mean, running_var, momentum, eps, training, group, n_global)


# ----------------------------------------------------------------------------------------------
# completion scorer
# ----------------------------------------------------------------------------------------------
def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class TransEMarginFn(torch.autograd.Function):
    """loss[l] of include/jmac_b200.h:jmac_transe_margin_fwd for n_layers (ent, rel) table pairs.
    Backward is an atomics-free segmented reduction over the radix-sorted occurrences of every table row
    (csrc/scorer.cu): the dense gradient tables are bit-identical from run to run."""

    @staticmethod
    def forward(ctx, h, r, t, batch_size: int, margin: float, *tables):
        lib = _lib.load()
        n_layers = len(tables) // 2
        ents = [_f32c(x) for x in tables[:n_layers]]
        rels = [_f32c(x) for x in tables[n_layers:]]
        _need_cuda(*ents, *rels)
        dev = ents[0].device
        d = ents[0].shape[1]
        if any(e.shape != ents[0].shape for e in ents) or any(r.shape != rels[0].shape for r in rels):
            raise ValueError("every layer's entity (relation) table must have the same shape: one index list addresses all of them")
        h, r, t = as_index(h, dev), as_index(r, dev), as_index(t, dev)
        M = int(h.numel())
        if int(r.numel()) != M or int(t.numel()) != M:
            raise RuntimeError("batch_h, batch_r, batch_t must have the same length")
        B = min(int(batch_size), M)
        if M % B != 0 or M // B < 2:
            # same failure the reference hits in score.view(-1, batch_size) (src/jmac_model.py:358,367)
            raise RuntimeError(f"shape '[-1, {B}]' is invalid for input of size {M - B}")
        _lib.raise_if_flagged()
        loss = torch.empty(n_layers, dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad)
        record = None
        if need_grad:
            record = torch.empty(int(lib.jmac_transe_record_bytes(M, B, n_layers, d)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.jmac_transe_margin_fwd(_ptr_array(ents), _ptr_array(rels), n_layers, d, ptr(h), ptr(r), ptr(t), M, B,
                                             ents[0].shape[0], rels[0].shape[0], float(margin), None, ptr(loss), ptr(record), stream_ptr()),
                  "jmac_transe_margin_fwd")
        if need_grad:
            ctx.save_for_backward(h, r, t, record)
        ctx.meta = (n_layers, d, M, B, tuple(ents[0].shape), tuple(rels[0].shape), dev)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        lib = _lib.load()
        n_layers, d, M, B, eshape, rshape, dev = ctx.meta
        h, r, t, record = ctx.saved_tensors
        gloss = _f32c(gloss)
        d_ents = [torch.zeros(eshape, dtype=torch.float32, device=dev) for _ in range(n_layers)]
        d_rels = [torch.zeros(rshape, dtype=torch.float32, device=dev) for _ in range(n_layers)]
        ws_bytes = int(lib.jmac_transe_bwd_workspace_bytes(M, B, n_layers, d))
        if ws_bytes == 0:
            raise _lib.JmacError("batch too large for the 30-bit occurrence words of the scorer's backward (2*M + B < 2^30)")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.jmac_transe_margin_bwd(n_layers, d, ptr(h), ptr(r), ptr(t), M, B, eshape[0], rshape[0], ptr(record), ptr(gloss),
                                             _ptr_array(d_ents), _ptr_array(d_rels), ptr(ws), ws_bytes, stream_ptr()),
                  "jmac_transe_margin_bwd")
        return (None, None, None, None, None, *d_ents, *d_rels)


def transe_margin_loss(ents: List[torch.Tensor], rels: List[torch.Tensor], h, r, t, batch_size: int, margin: float) -> torch.Tensor:
    """Per-layer losses [n_layers] (sum them as src/jmac_model.py:378 does)."""
    return TransEMarginFn.apply(h, r, t, batch_size, margin, *ents, *rels)


class PairCosFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, ia, ib):
        lib = _lib.load()
        A, B = _f32c(A), _f32c(B)
        _need_cuda(A, B)
        dev = A.device
        ia, ib = as_index(ia, dev).reshape(-1), as_index(ib, dev).reshape(-1)
        M = int(ia.numel())
        _lib.raise_if_flagged()
        cosv = torch.empty(M, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.jmac_pair_cos_fwd(ptr(A), ptr(B), A.shape[1], ptr(ia), ptr(ib), M, A.shape[0], B.shape[0], ptr(cosv), stream_ptr()),
                  "jmac_pair_cos_fwd")
        ctx.save_for_backward(A, B, ia, ib)
        return cosv

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        A, B, ia, ib = ctx.saved_tensors
        g = _f32c(g)
        dA = torch.zeros_like(A)
        dB = torch.zeros_like(B)
        with torch.cuda.device(A.device):
            check(lib.jmac_pair_cos_bwd(ptr(A), ptr(B), A.shape[1], ptr(ia), ptr(ib), int(ia.numel()), A.shape[0], B.shape[0], ptr(g), ptr(dA),
                                        ptr(dB), stream_ptr()),
                  "jmac_pair_cos_bwd")
        return dA, dB, None, None


def pair_cos(A, B, ia, ib) -> torch.Tensor:
    return PairCosFn.apply(A, B, ia, ib)


# ----------------------------------------------------------------------------------------------
# link prediction
# ----------------------------------------------------------------------------------------------
def l1_dist(ents, rels, e_idx, r_idx, pred_head=False) -> torch.Tensor:
    lib = _lib.load()
    ents = [_f32c(x.detach()) for x in ents]
    rels = [_f32c(x.detach()) for x in rels]
    _need_cuda(*ents, *rels)
    dev = ents[0].device
    e_idx, r_idx = as_index(e_idx, dev), as_index(r_idx, dev)
    Bq, N, d = int(e_idx.numel()), ents[0].shape[0], ents[0].shape[1]
    dist = torch.empty(Bq, N, dtype=torch.float32, device=dev)
    if Bq == 0:
        return dist
    _lib.raise_if_flagged()
    with torch.cuda.device(dev):
        check(lib.jmac_l1_dist(_ptr_array(ents), _ptr_array(rels), len(ents), d, ptr(e_idx), ptr(r_idx), Bq, N, rels[0].shape[0], int(pred_head),
                               ptr(dist), stream_ptr()), "jmac_l1_dist")
    return dist


def l1_rank(ents, rels, e_idx, r_idx, t_idx, filt_ptr=None, filt_idx=None, pred_head=False) -> torch.Tensor:
    """1-based (filtered) rank of each target by counting; int32 [B]."""
    lib = _lib.load()
    ents = [_f32c(x.detach()) for x in ents]
    rels = [_f32c(x.detach()) for x in rels]
    _need_cuda(*ents, *rels)
    dev = ents[0].device
    e_idx, r_idx, t_idx = as_index(e_idx, dev), as_index(r_idx, dev), as_index(t_idx, dev)
    Bq, N, d = int(e_idx.numel()), ents[0].shape[0], ents[0].shape[1]
    rank = torch.empty(Bq, dtype=torch.int32, device=dev)
    if Bq == 0:
        return rank
    if filt_ptr is not None:
        filt_ptr, filt_idx = as_index(filt_ptr, dev), as_index(filt_idx, dev)
    _lib.raise_if_flagged()
    with torch.cuda.device(dev):
        check(lib.jmac_l1_rank(_ptr_array(ents), _ptr_array(rels), len(ents), d, ptr(e_idx), ptr(r_idx), ptr(t_idx), Bq, N, rels[0].shape[0],
                               int(pred_head), ptr(filt_ptr), ptr(filt_idx), ptr(rank), stream_ptr()), "jmac_l1_rank")
    return rank
