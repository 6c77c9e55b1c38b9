"""Destination-sharded execution of the relation-aware GNN encoder and data-parallel completion scoring on the
GPUs of one node (one process per GPU, ``torch.distributed`` over NCCL / NVLink 5).

The reference is single-device; this is new design (BASELINE.json north_star):
  * nodes are split into contiguous ranges balanced by IN-EDGE count (hubs!), every rank owns the
    CSR rows of its destinations and the matching slice of the entity table;
  * per layer forward: all-gather of the X rows (first layer: the replicated input is re-laid-out locally,
    no communication), Q = X W_bot recomputed on the gathered table (a 1.5 ms GEMM instead of a second
    2 GB all-gather); edge kernels run locally on a bipartite plan (local destinations x gathered sources);
  * BatchNorm couples all shards: all-reduce of the 2d per-feature sums (functional.BNTanhFn);
  * backward: dXs = dQ W_bot^T + dXe are partial sums over the local edges for ALL source rows ->
    reduce-scatter to the owners (the autograd transpose of the all-gather);
  * relation-table, layer-weight and BN-affine gradients are all-reduced once per step.

Overlap (``ShardedAttnEdgeFn``, the CUDA path): the gathered table is laid out in ``chunks`` row groups
(``Shard.chunks``), each exchanged by its own collective on NCCL's stream.  Forward: the Q GEMM of chunk c
starts as soon as chunk c has landed, while chunks c+1.. are still in flight (and the local P GEMM runs
under the first chunk).  Backward: the reduce-scatter of chunk c is issued as soon as its
``dQ W_bot^T + dXe`` GEMM is done and proceeds under the remaining chunk GEMMs, the relation pass, the local
input-gradient GEMM and the three weight-gradient GEMMs; the compute stream waits for it only at the very end
of the node's backward.  Collectives are real exchange steps of the algorithm; nothing is invented for link
prediction / alignment, which shard by independent queries / rows.

Completion scorer: ``sharded_transe_margin_loss`` shards the B positives (with all of their
negatives) over the ranks; every rank scores its share against the full tables (replicated parameters or
``GatherRowsFn`` outputs of row-sharded encoder outputs), the loss is all-reduced, the table gradients go back
through the reduce-scatter transpose of the gather (sharded tables) or ``all_reduce_grads`` (replicated ones).

``ops`` is the compute backend: ``jmac_b200.functional`` (CUDA) by default.  CPU tests inject a
torch implementation to check the orchestration with the gloo backend (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# partitioning
# ----------------------------------------------------------------------------------------------
def partition_by_in_edges(row: torch.Tensor, num_ent: int, world: int) -> List[int]:
    """Boundaries b[0]=0 <= ... <= b[world]=num_ent of contiguous node ranges with (nearly) equal
    numbers of in-edges; every range is non-empty when num_ent >= world."""
    deg = torch.bincount(row.reshape(-1).to(torch.int64), minlength=num_ent)
    cum = torch.cumsum(deg, 0)
    total = int(cum[-1].item()) if num_ent > 0 else 0
    bounds = [0]
    for g in range(1, world):
        target = (total * g + world - 1) // world
        b = int(torch.searchsorted(cum, torch.tensor([target], device=cum.device, dtype=cum.dtype)).item()) + 1 if total > 0 else 0
        b = max(b, bounds[-1] + 1)                       # non-empty
        b = min(b, num_ent - (world - g))                # leave room for the remaining ranks
        bounds.append(b)
    bounds.append(num_ent)
    return bounds


@dataclass
class Shard:
    rank: int
    world: int
    bounds: List[int]
    n_max: int              # rows per rank in gathered tables (max shard size rounded up to a multiple of chunks)
    lo: int
    hi: int
    edge_index: torch.Tensor    # [2, E_loc]: row = local destination, col = row in the gathered table
    edge_type: torch.Tensor
    n_edges_global: int
    chunks: int = 1             # row groups of the gathered layout, each exchanged by its own collective
    peer: object = None         # jmac_b200.peer.PeerExchange: copy-engine exchange over NVLink peer memory instead of NCCL

    @property
    def n_loc(self):
        return self.hi - self.lo

    @property
    def n_c(self):
        return self.n_max // self.chunks

    @property
    def n_src(self):
        return self.world * self.n_max

    @property
    def n_total(self):
        return self.bounds[-1]

    def chunk_rows(self, c):
        """Row range of chunk c in the gathered [world * n_max, d] table."""
        w = self.world * self.n_c
        return c * w, (c + 1) * w

    def gathered_row(self, owner, local):
        """Row of local row ``local`` of rank ``owner`` in the gathered table: chunk-major, then owner, then row
        (with chunks == 1 this is owner * n_max + local).  Works on ints and on tensors."""
        c = local // self.n_c
        return c * (self.world * self.n_c) + owner * self.n_c + (local - c * self.n_c)

    def to_gathered(self, idx: torch.Tensor) -> torch.Tensor:
        """Global entity ids -> rows of the gathered table (so that index lists can address it directly)."""
        bt = torch.tensor(self.bounds, device=idx.device, dtype=idx.dtype)
        owner = torch.bucketize(idx, bt[1:], right=True)
        return self.gathered_row(owner, idx - bt[owner])

    def gather_map(self, device):
        """Row of the replicated [N,d] table held at each row of the gathered layout (padding rows point at
        row 0; they are never referenced by an edge)."""
        m = getattr(self, "_gmap", None)
        if m is None or m.device != torch.device(device):
            m = torch.zeros(self.n_src, dtype=torch.long)
            for g in range(self.world):
                ng = self.bounds[g + 1] - self.bounds[g]
                loc = torch.arange(ng)
                m[self.gathered_row(g, loc)] = self.bounds[g] + loc
            m = m.to(device)
            object.__setattr__(self, "_gmap", m)
        return m


def build_shard(edge_index: torch.Tensor, edge_type: torch.Tensor, num_ent: int, rank: int, world: int,
                bounds: Optional[Sequence[int]] = None, chunks: int = 1) -> Shard:
    row, col = edge_index[0], edge_index[1]
    bounds = list(bounds) if bounds is not None else partition_by_in_edges(row, num_ent, world)
    chunks = max(1, int(chunks))
    n_raw = max(bounds[g + 1] - bounds[g] for g in range(world))
    n_max = (n_raw + chunks - 1) // chunks * chunks
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = (row >= lo) & (row < hi)
    r, c, t = row[mine], col[mine], edge_type[mine]
    bt = torch.tensor(bounds, device=c.device, dtype=c.dtype)
    owner = torch.bucketize(c, bt[1:], right=True)                     # owner rank of every source
    sh = Shard(rank, world, bounds, n_max, lo, hi, None, t.contiguous(), int(row.numel()), chunks)
    c_g = sh.gathered_row(owner, c - bt[owner])                         # row in the gathered table
    sh.edge_index = torch.stack([r - lo, c_g]).contiguous()
    return sh


# ----------------------------------------------------------------------------------------------
# collectives with their autograd transposes
# ----------------------------------------------------------------------------------------------
def _backend(group) -> str:
    return dist.get_backend(group)


import os as _os
# diagnostics (profiles/multi_gpu_overlap_r02.md): "nocomm" skips the exchanges (pure per-rank compute time, results meaningless),
# "serial" waits for every collective right after issuing it (no overlap)
_DIAG = _os.environ.get("JMAC_DIST_DIAG", "")


def _pad_rows(x_loc: torch.Tensor, n_max: int) -> torch.Tensor:
    if x_loc.shape[0] < n_max:
        x_loc = torch.cat([x_loc, x_loc.new_zeros(n_max - x_loc.shape[0], x_loc.shape[1])])
    return x_loc.contiguous()


def gather_rows(x_loc: torch.Tensor, shard: Shard, group, async_op: bool = False):
    """[n_loc, d] on every rank -> the gathered [world * n_max, d] table (layout: Shard.gathered_row).
    ``async_op``: returns (table, [work per chunk]) without waiting (NCCL: the collectives run on the
    communicator's stream; ``work.wait()`` makes the current stream wait for chunk c only)."""
    d = x_loc.shape[1]
    buf = _pad_rows(x_loc, shard.n_max)
    out = x_loc.new_empty(shard.n_src, d)
    works = []
    for c in range(shard.chunks):
        r0, r1 = shard.chunk_rows(c)
        src = buf[c * shard.n_c:(c + 1) * shard.n_c]
        if _DIAG == "nocomm":
            works.append(None)
        elif _backend(group) == "nccl":
            w = dist.all_gather_into_tensor(out[r0:r1], src, group=group, async_op=async_op)
            if _DIAG == "serial" and w is not None:
                w.wait()
                w = None
            works.append(w)
        else:
            dist.all_gather(list(out[r0:r1].view(shard.world, shard.n_c, d).unbind(0)), src.contiguous(), group=group)
            works.append(None)
    return (out, works) if async_op else out


def scatter_sum_chunk(g_chunk: torch.Tensor, out_chunk: torch.Tensor, shard: Shard, group, async_op: bool = False):
    """Sum over ranks of one chunk [world * n_c, d] of partial rows; this rank keeps its n_c rows (``out_chunk``)."""
    if _DIAG == "nocomm":
        return None
    if _backend(group) == "nccl":
        w = dist.reduce_scatter_tensor(out_chunk, g_chunk, group=group, async_op=async_op)
        if _DIAG == "serial" and w is not None:
            w.wait()
            w = None
        return w
    tmp = g_chunk.clone()                      # gloo has no reduce-scatter
    dist.all_reduce(tmp, group=group)
    out_chunk.copy_(tmp.view(shard.world, shard.n_c, -1)[shard.rank])
    return None


def scatter_sum_rows(g_full: torch.Tensor, shard: Shard, group) -> torch.Tensor:
    """Transpose of gather_rows: sum over ranks of [world * n_max, d], keep this rank's rows."""
    g_full = g_full.contiguous()
    out = g_full.new_empty(shard.n_max, g_full.shape[1])
    for c in range(shard.chunks):
        r0, r1 = shard.chunk_rows(c)
        scatter_sum_chunk(g_full[r0:r1], out[c * shard.n_c:(c + 1) * shard.n_c], shard, group)
    return out[: shard.n_loc].contiguous()


class GatherRowsFn(torch.autograd.Function):
    """x_loc [n_loc,d] -> gathered [world*n_max, d]; backward = reduce-scatter (sum) of the partials."""

    @staticmethod
    def forward(ctx, x_loc, shard, group, x_full=None):
        ctx.shard, ctx.group = shard, group
        if x_full is not None:
            # the input table is replicated on every rank (first layer): build the gathered layout locally,
            # no communication; the gradient still has to be reduce-scattered
            return x_full.detach().index_select(0, shard.gather_map(x_full.device))
        return gather_rows(x_loc, shard, group)

    @staticmethod
    def backward(ctx, g):
        return scatter_sum_rows(g, ctx.shard, ctx.group), None, None, None


class AllReduceSumFn(torch.autograd.Function):
    """y = sum over ranks of x.  Backward is the identity: the job's loss is the SUM of the ranks' terms and every
    rank back-propagates its own term (the partial gradients meet in reduce-scatters / all_reduce_grads)."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.clone()
        dist.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


# ----------------------------------------------------------------------------------------------
# sharded layer / encoder
# ----------------------------------------------------------------------------------------------
def _default_ops():
    from . import functional as JF
    return JF


class _EventWait:
    """``work.wait()`` look-alike over CUDA events: the current stream waits for them."""

    def __init__(self, events):
        self.events = events

    def wait(self):
        cur = torch.cuda.current_stream()
        for ev in self.events:
            cur.wait_event(ev)


class ShardedAttnEdgeFn(torch.autograd.Function):
    """``functional.AttnEdgeFn`` for one destination shard with the exchange fused in and overlapped (see the module
    docstring): forward = chunked all-gather of X || P GEMM, then per chunk Q GEMM as it lands, edge kernel;
    backward = destination pass, source pass, per chunk (dQ W_bot^T + dXe GEMM -> asynchronous reduce-scatter) ||
    relation pass, local input-gradient GEMM, weight-gradient GEMMs; wait; dX_loc = dP W_top^T + dT + scattered sum."""

    @staticmethod
    def forward(ctx, plan, x_loc, x_full, W_att, Rp, a_att, slope, shard, group):
        from . import functional as JF
        from ._lib import check, ptr, stream_ptr, load
        lib = load()
        x_loc, W_att, Rp, a = JF._f32c(x_loc), JF._f32c(W_att), JF._f32c(Rp), JF._f32c(a_att.reshape(-1))
        n_loc, d = x_loc.shape
        dev = x_loc.device
        Wt, Wb = W_att[:d], W_att[d:]
        peer = shard.peer if _DIAG != "nocomm" else None
        if x_full is not None:
            Xf, works = x_full.detach().index_select(0, shard.gather_map(dev)), [None] * shard.chunks
        elif peer is not None:
            Xf = torch.empty(shard.n_src, d, dtype=torch.float32, device=dev)
            works = [_EventWait(evs) for evs in peer.all_gather(x_loc.detach(), Xf)]    # copy-engine pulls, no SM involved
            if _DIAG == "serial":
                for w in works:
                    w.wait()
        else:
            Xf, works = gather_rows(x_loc.detach(), shard, group, async_op=True)
        P = JF.gemm_nn(x_loc, Wt)                                  # runs under the first chunk of the all-gather
        S = JF.small_mm(Rp, Wb)
        wb_hi, wb_lo = JF._split_weight(Wb, True)                   # split once, used by every chunk
        Q = torch.empty_like(Xf)
        for c in range(shard.chunks):
            if works[c] is not None:
                works[c].wait()                                     # the compute stream waits for chunk c only
            r0, r1 = shard.chunk_rows(c)
            JF._gemm_presplit(Xf[r0:r1], wb_hi, wb_lo, out=Q[r0:r1])
        agg, T = torch.empty_like(x_loc), torch.empty_like(x_loc)
        lse = torch.empty(n_loc, dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad)
        erec = torch.empty(max(plan.E, 1) * int(lib.jmac_edge_record_bytes(d)), dtype=torch.uint8, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            check(lib.jmac_edge_fwd_fused(plan.handle, ptr(P), ptr(Q), ptr(Xf), ptr(S), ptr(Rp), ptr(a), float(slope), d, ptr(agg), ptr(lse),
                                          ptr(x_loc), ptr(T), ptr(erec), stream_ptr()), "jmac_edge_fwd")
        ctx.plan, ctx.slope, ctx.a_shape, ctx.shard, ctx.group = plan, float(slope), a_att.shape, shard, group
        if need_grad:
            ctx.save_for_backward(x_loc, Xf, W_att, Rp, a, P, Q, S, agg, lse, erec)
        return T

    @staticmethod
    def backward(ctx, g):
        from . import functional as JF
        from ._lib import check, ptr, stream_ptr, load
        lib = load()
        x_loc, Xf, W_att, Rp, a, P, Q, S, agg, lse, erec = ctx.saved_tensors
        plan, slope, shard, group = ctx.plan, ctx.slope, ctx.shard, ctx.group
        n_loc, d = x_loc.shape
        dev = x_loc.device
        G = JF._f32c(g)
        Wt, Wb = W_att[:d], W_att[d:]
        peer = shard.peer if _DIAG != "nocomm" else None
        dP, dQ = torch.empty_like(x_loc), torch.empty_like(Xf)
        dXe = peer.partial_table() if peer is not None else torch.empty_like(Xf)      # peers pull their rows straight out of it
        dRp, dS = torch.empty_like(Rp), torch.empty_like(S)
        acc2 = torch.zeros(2, d, dtype=torch.float32, device=dev)     # [0] = d a_att, [1] = column sums of G over this shard
        da, gsum = acc2[0], acc2[1]
        with torch.cuda.device(dev):
            st = stream_ptr()
            check(lib.jmac_edge_bwd_dst(plan.handle, ptr(P), ptr(Xf), ptr(Rp), ptr(a), slope, d, ptr(agg), ptr(lse), ptr(G), ptr(dP),
                                        ptr(erec), ptr(da), ptr(gsum), st), "jmac_edge_bwd_dst")
            check(lib.jmac_edge_bwd_src(plan.handle, ptr(G), ptr(erec), ptr(a), slope, d, ptr(Q), ptr(dXe), ptr(dQ), ptr(da), st),
                  "jmac_edge_bwd_src")
        # the partial input gradient of every gathered row, chunk by chunk, each chunk handed to NCCL as soon as it exists
        wb_hi, wb_lo = JF._split_weight(Wb, False)
        scattered = torch.empty(shard.n_max, d, dtype=torch.float32, device=dev) if peer is None else None
        works = []
        for c in range(shard.chunks):
            r0, r1 = shard.chunk_rows(c)
            JF._gemm_presplit(dQ[r0:r1], wb_hi, wb_lo, add1=dXe[r0:r1], out=dXe[r0:r1])      # dQ W_bot^T + dXe, in place
            if peer is not None:
                peer.scatter_chunk(c)                                                       # handshake + copy-engine pulls of my rows
                if _DIAG == "serial":
                    peer.wait_pulls()
            else:
                works.append(scatter_sum_chunk(dXe[r0:r1], scattered[c * shard.n_c:(c + 1) * shard.n_c], shard, group, async_op=True))
        # ... under which the rest of the node's backward runs on the compute stream
        with torch.cuda.device(dev):
            check(lib.jmac_edge_bwd_rel(plan.handle, ptr(G), ptr(erec), ptr(a), slope, d, ptr(S), ptr(dRp), ptr(dS), ptr(da), stream_ptr()),
                  "jmac_edge_bwd_rel")
        dRp[-1] -= gsum                                       # T = agg + x_loc - Rp[-1]
        JF.small_mm_nt_acc(dS, Wb, dRp)                       # S = Rp Wb: dRp += dS Wb^T
        dXd = JF.gemm_nt(dP, Wt, add1=G)                      # dP W_top^T + dT
        dW = None
        if ctx.needs_input_grad[3]:
            dW = torch.empty_like(W_att)
            dW[:d] = JF.gemm_tn(x_loc, dP)
            dWb = JF.gemm_tn(Xf, dQ)
            dWb += JF.small_mm_tn(Rp, dS)
            dW[d:] = dWb
        if peer is not None:
            peer.scatter_finish(dXd)                          # waits for the pulls, then dXd += sum over ranks (one kernel)
        else:
            for w in works:
                if w is not None:
                    w.wait()
            dXd += scattered[:n_loc]
        return None, dXd, None, dW, dRp, da.view(ctx.a_shape), None, None, None


def _fused_eligible(ops, x_loc, shard, d) -> bool:
    if getattr(ops, "attn_edge_eligible", None) is None or not x_loc.is_cuda:
        return False
    chunk_rows = shard.world * shard.n_c
    return ops.attn_edge_eligible(x_loc, d) and ops._tc_eligible(chunk_rows, d, d) and ops._tn_eligible(shard.n_src, d, d)


def sharded_layer_forward(layer, x_loc, rel_emb, shard: Shard, plan, group, ops=None, x_full=None):
    """RelationAwareLayer.forward (src/jmac_model.py:33-53) for the destination range of this rank.
    ``plan`` is the bipartite plan of shard.edge_index (n_loc destinations x world*n_max sources)."""
    ops = ops or _default_ops()
    d = layer.in_channels
    if layer.comp_op not in ("sub", "mult"):
        raise NotImplementedError
    Rp = layer.relation_transform(rel_emb)
    x_loc = x_loc.contiguous()
    mm = getattr(ops, "node_mm", torch.mm)
    if layer.comp_op == "mult":
        # src/jmac_model.py:63-64 under destination sharding: the per-edge tables are built from the gathered source table for the
        # local destinations only (model.RelationAwareLayer.mult_edge_phase); the exchange is the same all-gather / reduce-scatter
        Xf = GatherRowsFn.apply(x_loc, shard, group, x_full)
        T = layer.mult_edge_phase(x_loc, Xf, Rp, shard.edge_index, shard.edge_type, ops)
    elif _fused_eligible(ops, x_loc, shard, d):
        T = ShardedAttnEdgeFn.apply(plan, x_loc, x_full, layer.w_att, Rp, layer.a_att, layer.leaky_slope, shard, group)
    else:
        # small shards (library GEMMs) and the CPU stand-ins of the gloo tests: the same algebra through autograd --
        # only X is gathered, Q = X W_bot is recomputed on the gathered table, autograd sums dXe + dQ W_bot^T on the
        # partials BEFORE the single reduce-scatter (linearity)
        Xf = GatherRowsFn.apply(x_loc, shard, group, x_full)
        P = mm(x_loc, layer.w_att[:d])
        S = torch.mm(Rp, layer.w_att[d:])
        Qf = mm(Xf, layer.w_att[d:])
        T = ops.edge_phase(plan, P, Qf, Xf, S, Rp, layer.a_att, layer.leaky_slope, x_loc)     # self-loop term fused
    pre = mm(T, layer.gcn_weight * 0.5)
    training = layer.training or not layer.bn.track_running_stats
    if layer.training and layer.bn.track_running_stats:
        layer.bn.num_batches_tracked += 1
    momentum = layer.bn.momentum if layer.bn.momentum is not None else 0.1
    return ops.bn_tanh(pre, layer.bn.weight, layer.bn.bias, layer.bn.running_mean, layer.bn.running_var, momentum, layer.bn.eps, training, group,
                       shard.n_total)


def all_reduce_grads(params, group):
    """Sum the gradients of replicated parameters (relation table, layer weights, BN affine) over the
    destination shards, in one flat buffer."""
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, group=group)
    off = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def default_chunks(group, n_rows_per_rank: int) -> int:
    """Exchange chunks per layer: 4 on NCCL when every chunk still feeds the tensor-core GEMMs (>= 4096 rows), else 1."""
    if _backend(group) != "nccl":
        return 1
    if _os.environ.get("JMAC_DIST_CHUNKS"):
        return max(1, int(_os.environ["JMAC_DIST_CHUNKS"]))
    world = dist.get_world_size(group)
    for c in ((2,) if world > 4 else (4, 2)):           # 8 GPUs: 35.6 ms with 2 chunks, 36.4-36.6 ms with 4 or 8 (profiles/multi_gpu_overlap_r02.md)
        if n_rows_per_rank // c * world >= 8192:
            return c
    return 1


class ShardedEncoder:
    """Stack of RelationAwareLayers over one destination-sharded graph (bench.py --gpus N and
    tests/test_gpu_dist.py)."""

    def __init__(self, layers, edge_index, edge_type, num_ent: int, num_rel: int, group=None, ops=None, bounds=None, chunks=None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.layers = list(layers)
        self.ops = ops or _default_ops()
        if chunks is None:
            chunks = default_chunks(self.group, (num_ent + self.world - 1) // self.world)
        self.shard = build_shard(edge_index, edge_type, num_ent, self.rank, self.world, bounds, chunks)
        self.plan = self.ops.GraphPlan(self.shard.edge_index, self.shard.edge_type, self.shard.n_loc, num_rel + 1, n_src=self.shard.n_src)
        self.replicated = [p for l in self.layers for p in l.parameters()]
        self._x_loc = None
        self._setup_peer()

    def _setup_peer(self):
        """Copy-engine exchange over peer memory (jmac_b200/peer.py) when every rank can set it up; NCCL collectives otherwise.
        Default: up to 4 ranks -- measured on the C4 graph (profiles/multi_gpu_overlap_r02.md): 73.9 vs 77.4 ms at 2 GPUs, 46.6 vs 49.6 ms
        at 4, but 39.7 vs 36.6 ms at 8, where seven concurrent copy-engine pulls per rank no longer reach NCCL's transfer rate.
        ``JMAC_DIST_PEER=0`` / ``=1`` forces NCCL / peer memory.  The decision is collective: either every rank uses it or none."""
        want = _os.environ.get("JMAC_DIST_PEER", "")
        if _backend(self.group) != "nccl" or self.world < 2 or want == "0" or (want != "1" and self.world > 4):
            return
        d = self.layers[0].in_channels
        dev = self.shard.edge_index.device
        ops, sh = self.ops, self.shard
        if getattr(ops, "attn_edge_eligible", None) is None or not dev.type == "cuda":
            return
        mine = (all(l.in_channels == d and l.comp_op == "sub" for l in self.layers) and ops._tc_eligible(sh.n_loc, d, d) and ops._tn_eligible(sh.n_loc, d, d)
                and ops._tc_eligible(sh.world * sh.n_c, d, d) and ops._tn_eligible(sh.n_src, d, d))
        # The handshakes park cuStreamWaitValue32 waits on the copy streams.  With the default 8 hardware work queues a copy stream can share
        # a queue with the compute stream, and a parked wait would then hold back the very kernels the peer is waiting for: the exchange is
        # only used when the process was started with enough queues (the variable is read when the CUDA context is created).
        if int(_os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "8") or 8) < 16 and want != "1":
            if mine and self.rank == 0:
                import warnings
                warnings.warn("peer-memory exchange not used: start the process with CUDA_DEVICE_MAX_CONNECTIONS=32 (NCCL collectives instead)")
            mine = False
        vote = torch.tensor([1.0 if mine else 0.0], device=dev)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN, group=self.group)
        if float(vote.item()) == 0.0:
            return                                           # some rank would take the composed path: NCCL for everybody
        from .peer import PeerExchange
        try:
            self.shard.peer = PeerExchange(self.shard, d, self.group, dev)
        except RuntimeError as e:
            if self.rank == 0:
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({e}); using NCCL collectives")
            self.shard.peer = None

    def close(self):
        """Release the peer-memory buffers (collective in effect: call it on every rank, after the last step has completed)."""
        peer, self.shard.peer = self.shard.peer, None
        if peer is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)           # nobody may still be pulling from the buffers freed below
            peer.close(collective=True)

    def forward(self, x_loc, rel_emb, x_full=None):
        """``x_full``: the replicated input table, if every rank holds it (saves the first all-gather)."""
        h = x_loc
        for i, layer in enumerate(self.layers):
            h = sharded_layer_forward(layer, h, rel_emb, self.shard, self.plan, self.group, self.ops, x_full if i == 0 else None)
        return h

    def step(self, x_full, rel_emb, gout_full):
        """One forward+backward over the whole (sharded) graph.  ``x_full``/``gout_full`` are the
        replicated [N,d] input table / upstream gradient; each rank works on its own rows (views, re-taken every call)."""
        sh = self.shard
        x_loc = x_full.detach()[sh.lo:sh.hi].requires_grad_(True)
        self._x_loc = x_loc                                   # the leaf whose .grad holds this rank's rows of dL/dX
        for p in self.replicated + [rel_emb]:
            p.grad = None
        out = self.forward(x_loc, rel_emb, x_full)
        out.backward(gout_full[sh.lo:sh.hi])
        all_reduce_grads(self.replicated + [rel_emb], self.group)
        return out


# ----------------------------------------------------------------------------------------------
# data-parallel completion scorer
# ----------------------------------------------------------------------------------------------
def positive_share(batch_size: int, n_triples: int, rank: int, world: int):
    """This rank's positives [b0, b1) of the B = min(batch_size, M) positives and the flat positions of their
    (1+K) triples in the reference layout (train.py:347-352: position k*B + b), re-laid-out as k*(b1-b0) + (b-b0)."""
    B = min(int(batch_size), int(n_triples))
    K1 = n_triples // B
    per = (B + world - 1) // world
    b0 = min(rank * per, B)
    b1 = min(b0 + per, B)
    sel = (torch.arange(K1)[:, None] * B + torch.arange(b0, b1)[None, :]).reshape(-1)
    return B, b0, b1, sel


def sharded_transe_margin_loss(ents, rels, h, r, t, batch_size: int, margin: float, group, ops=None) -> torch.Tensor:
    """Per-layer losses [n_layers] of JMAC.completion_loss's scorer (src/jmac_model.py:345-374), the positives sharded
    over the ranks of ``group``.  ``ents`` / ``rels`` are FULL tables on every rank (replicated leaves, or the output of
    ``GatherRowsFn`` for row-sharded encoder outputs); h, r, t the full (1+K)*B index lists, identical on every rank.
    Value: identical on every rank (all-reduced).  Backward: every rank back-propagates its own share; gradients of
    gathered tables are reduce-scattered by the gather's transpose, replicated leaves need ``all_reduce_grads``."""
    ops = ops or _default_ops()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = ents[0].device
    h, r, t = (ops.as_index(x, dev).reshape(-1) for x in (h, r, t))
    M = int(h.numel())
    if int(r.numel()) != M or int(t.numel()) != M:
        raise RuntimeError("batch_h, batch_r, batch_t must have the same length")
    B, b0, b1, sel = positive_share(batch_size, M, rank, world)
    if M % B != 0 or M // B < 2:
        raise RuntimeError(f"shape '[-1, {B}]' is invalid for input of size {M - B}")
    if b1 > b0:
        sel = sel.to(dev)
        local = ops.transe_margin_loss(list(ents), list(rels), h[sel], r[sel], t[sel], b1 - b0, margin)
        # loss = mean over all B*K hinge terms + margin = sum_ranks (B_r / B) * (mean_r + margin)
        local = local * ((b1 - b0) / B)
    else:
        local = sum(e.sum() * 0 for e in ents).expand(len(ents)) + torch.zeros(len(ents), dtype=ents[0].dtype, device=dev)
    return AllReduceSumFn.apply(local, group)
