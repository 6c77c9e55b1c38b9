"""Destination-sharded execution of the relation-aware GNN encoder on the GPUs of one node
(one process per GPU, ``torch.distributed`` over NCCL / NVLink 5).

The reference is single-device; this is new design (BASELINE.json north_star):
  * nodes are split into contiguous ranges balanced by IN-EDGE count (hubs!), every rank owns the
    CSR rows of its destinations and the matching slice of the entity table;
  * per layer forward: local GEMMs P = X_loc W_top, Q_loc = X_loc W_bot, then a halo all-gather of
    the Q and X rows (first version: the full tables, 2 x 2 GB at C4 -- NVSwitch gives every GPU full
    bandwidth to every peer, so the cost is bytes / link bandwidth, not topology);
  * edge kernels run locally on a bipartite plan (local destinations x gathered sources);
  * BatchNorm couples all shards: all-reduce of the 2d+1 per-feature sums (functional.BNTanhFn);
  * backward: the source-pass gradients dXe, dQ are partial sums over the local edges for ALL source
    rows -> reduce-scatter to the owners (the autograd transpose of the all-gather);
  * relation-table, layer-weight and BN-affine gradients are all-reduced once per step.
Collectives are real exchange steps of the algorithm; nothing is invented for the scorers, which
shard by independent triples / rows.

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
    n_max: int              # rows per rank in gathered tables (max shard size)
    lo: int
    hi: int
    edge_index: torch.Tensor    # [2, E_loc]: row = local destination, col = row in the gathered table
    edge_type: torch.Tensor
    n_edges_global: int

    @property
    def n_loc(self):
        return self.hi - self.lo

    @property
    def n_src(self):
        return self.world * self.n_max

    def gather_map(self, device):
        """Row of the replicated [N,d] table held at each row of the gathered [world*n_max, d] layout
        (padding rows point at row 0; they are never referenced by an edge)."""
        m = getattr(self, "_gmap", None)
        if m is None or m.device != torch.device(device):
            parts = []
            for g in range(self.world):
                ng = self.bounds[g + 1] - self.bounds[g]
                parts.append(torch.arange(self.bounds[g], self.bounds[g + 1]))
                parts.append(torch.zeros(self.n_max - ng, dtype=torch.long))
            m = torch.cat(parts).to(device)
            object.__setattr__(self, "_gmap", m)
        return m


def build_shard(edge_index: torch.Tensor, edge_type: torch.Tensor, num_ent: int, rank: int, world: int,
                bounds: Optional[Sequence[int]] = None) -> Shard:
    row, col = edge_index[0], edge_index[1]
    bounds = list(bounds) if bounds is not None else partition_by_in_edges(row, num_ent, world)
    n_max = max(bounds[g + 1] - bounds[g] for g in range(world))
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = (row >= lo) & (row < hi)
    r, c, t = row[mine], col[mine], edge_type[mine]
    bt = torch.tensor(bounds, device=c.device, dtype=c.dtype)
    owner = torch.bucketize(c, bt[1:], right=True)                     # owner rank of every source
    c_g = owner * n_max + (c - bt[owner])                                # row in the [world * n_max] gathered table
    return Shard(rank, world, bounds, n_max, lo, hi, torch.stack([r - lo, c_g]).contiguous(), t.contiguous(), int(row.numel()))


# ----------------------------------------------------------------------------------------------
# collectives with their autograd transposes
# ----------------------------------------------------------------------------------------------
def _backend(group) -> str:
    return dist.get_backend(group)


def gather_rows(x_loc: torch.Tensor, shard: Shard, group) -> torch.Tensor:
    """[n_loc, d] on every rank -> [world * n_max, d] (rank g's rows at g*n_max .. g*n_max+n_g)."""
    d = x_loc.shape[1]
    buf = x_loc
    if shard.n_loc < shard.n_max:
        buf = torch.cat([x_loc, x_loc.new_zeros(shard.n_max - shard.n_loc, d)])
    buf = buf.contiguous()
    out = x_loc.new_empty(shard.world * shard.n_max, d)
    if _backend(group) == "nccl":
        dist.all_gather_into_tensor(out, buf, group=group)
    else:
        dist.all_gather(list(out.view(shard.world, shard.n_max, d).unbind(0)), buf, group=group)
    return out


def scatter_sum_rows(g_full: torch.Tensor, shard: Shard, group) -> torch.Tensor:
    """Transpose of gather_rows: sum over ranks of [world * n_max, d], keep this rank's rows."""
    d = g_full.shape[1]
    g_full = g_full.contiguous()
    if _backend(group) == "nccl":
        out = g_full.new_empty(shard.n_max, d)
        dist.reduce_scatter_tensor(out, g_full, group=group)
    else:                                    # gloo has no reduce-scatter
        tmp = g_full.clone()
        dist.all_reduce(tmp, group=group)
        out = tmp.view(shard.world, shard.n_max, d)[shard.rank]
    return out[: shard.n_loc].contiguous()


_PENDING = {}      # data_ptr(gathered tensor) -> outstanding NCCL work (the all-gather overlaps the local P GEMM)


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


def wait_gather(t: torch.Tensor):
    """Make the current stream wait for the all-gather that fills ``t`` (no-op if none is pending)."""
    w = _PENDING.pop(t.data_ptr(), None)
    if w is not None:
        w[0].wait()


# ----------------------------------------------------------------------------------------------
# sharded layer / encoder
# ----------------------------------------------------------------------------------------------
def _default_ops():
    from . import functional as JF
    return JF


def sharded_layer_forward(layer, x_loc, rel_emb, shard: Shard, plan, group, ops=None, x_full=None):
    """RelationAwareLayer.forward (src/jmac_model.py:33-53) for the destination range of this rank.
    ``plan`` is the bipartite plan of shard.edge_index (n_loc destinations x world*n_max sources)."""
    ops = ops or _default_ops()
    d = layer.in_channels
    if layer.comp_op != "sub":
        raise NotImplementedError
    Rp = layer.relation_transform(rel_emb)
    x_loc = x_loc.contiguous()
    mm = getattr(ops, "node_mm", torch.mm)
    # One exchange per layer and direction: only X is all-gathered; Q = X W_bot is recomputed on the gathered
    # table by every rank (a 1.5 ms GEMM instead of a second 2 GB all-gather), and in backward autograd sums
    # dXe + dQ W_bot^T on the partials BEFORE the single reduce-scatter (linearity).  The all-gather is
    # asynchronous and overlaps the local P GEMM.
    Xf = GatherRowsFn.apply(x_loc, shard, group, x_full)
    fused = getattr(ops, "attn_edge", None)
    if fused is not None and ops.attn_edge_eligible(x_loc, d) and ops.attn_edge_eligible(Xf, d):
        # one autograd node: dQ W_bot^T + dXe (the reduce-scattered partial) and dP W_top^T + dT come out of GEMM epilogues
        wait_gather(Xf)
        T = fused(plan, x_loc, layer.w_att, Rp, layer.a_att, layer.leaky_slope, Xf)
    else:
        P = mm(x_loc, layer.w_att[:d])
        S = torch.mm(Rp, layer.w_att[d:])
        wait_gather(Xf)
        Qf = mm(Xf, layer.w_att[d:])
        T = ops.edge_phase(plan, P, Qf, Xf, S, Rp, layer.a_att, layer.leaky_slope, x_loc)     # self-loop term fused
    pre = mm(T, layer.gcn_weight * 0.5)
    training = layer.training or not layer.bn.track_running_stats
    if layer.training and layer.bn.track_running_stats:
        layer.bn.num_batches_tracked += 1
    momentum = layer.bn.momentum if layer.bn.momentum is not None else 0.1
    return ops.bn_tanh(pre, layer.bn.weight, layer.bn.bias, layer.bn.running_mean, layer.bn.running_var, momentum, layer.bn.eps, training, group)


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


class ShardedEncoder:
    """Stack of RelationAwareLayers over one destination-sharded graph (bench.py --gpus N and
    tests/test_gpu_dist.py)."""

    def __init__(self, layers, edge_index, edge_type, num_ent: int, num_rel: int, group=None, ops=None, bounds=None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.layers = list(layers)
        self.ops = ops or _default_ops()
        self.shard = build_shard(edge_index, edge_type, num_ent, self.rank, self.world, bounds)
        self.plan = self.ops.GraphPlan(self.shard.edge_index, self.shard.edge_type, self.shard.n_loc, num_rel + 1, n_src=self.shard.n_src)
        self.replicated = [p for l in self.layers for p in l.parameters()]
        self._x_loc = None

    def forward(self, x_loc, rel_emb, x_full=None):
        """``x_full``: the replicated input table, if every rank holds it (saves the first all-gather)."""
        h = x_loc
        for i, layer in enumerate(self.layers):
            h = sharded_layer_forward(layer, h, rel_emb, self.shard, self.plan, self.group, self.ops, x_full if i == 0 else None)
        return h

    def step(self, x_full, rel_emb, gout_full):
        """One forward+backward over the whole (sharded) graph.  ``x_full``/``gout_full`` are the
        replicated [N,d] input table / upstream gradient; each rank works on its own rows."""
        sh = self.shard
        if self._x_loc is None:
            self._x_loc = x_full.detach()[sh.lo:sh.hi].clone().requires_grad_(True)
            self._gout_loc = gout_full[sh.lo:sh.hi].contiguous()
        x_loc = self._x_loc
        for p in self.replicated + [x_loc, rel_emb]:
            p.grad = None
        out = self.forward(x_loc, rel_emb, x_full)
        out.backward(self._gout_loc)
        all_reduce_grads(self.replicated + [rel_emb], self.group)
        return out
