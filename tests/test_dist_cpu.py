"""World-size-2/3 gloo tests (CPU) of the destination-sharded encoder's host logic: partitioning,
index remapping into the gathered table, the all-gather / reduce-scatter autograd pair, the
cross-shard BatchNorm reduction and the gradient all-reduce.  The CUDA ops are replaced by a torch
implementation of the same contracts (bipartite edge phase, BN+tanh with a group), and the sharded
result is compared with the single-process oracle in float64."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import jmac_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.clone()
        dist.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        y = g.clone()
        dist.all_reduce(y, group=ctx.group)
        return y, None


class CpuOps:
    """torch stand-ins honouring the contracts of jmac_b200.functional (test infrastructure)."""

    class GraphPlan:
        def __init__(self, edge_index, edge_type, num_ent, n_rel, chunk=0, n_src=None):
            self.ei, self.et, self.N, self.n_rel = edge_index, edge_type, num_ent, n_rel
            self.N_src = n_src if n_src is not None else num_ent

    @staticmethod
    def edge_phase(plan, P, Q, X, S, Rp, a_att, slope, Xself=None):
        row, col, t = plan.ei[0], plan.ei[1], plan.et
        z = P[row] + Q[col] - S[t]
        s = O.leaky_relu(z, slope) @ a_att.reshape(-1, 1)
        alpha = O.scatter_softmax(s, row, plan.N)
        deg = torch.bincount(row, minlength=plan.N).to(P.dtype)
        agg = O.scatter_sum(alpha * deg.sqrt()[row].view(-1, 1) * (X[col] - Rp[t]), row, plan.N)
        return agg if Xself is None else agg + Xself - Rp[-1]

    @staticmethod
    def bn_tanh(pre, gamma, beta, rm, rv, momentum, eps, training, group, n_global):
        n = torch.tensor([float(pre.shape[0])], dtype=pre.dtype)
        dist.all_reduce(n, group=group)
        assert int(n.item()) == n_global
        mean = _AllReduceSum.apply(pre.sum(0), group) / n
        var = _AllReduceSum.apply(((pre - mean) ** 2).sum(0), group) / n
        return torch.tanh((pre - mean) / torch.sqrt(var + eps) * gamma + beta)

    @staticmethod
    def as_index(x, device):
        return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).long()

    @staticmethod
    def transe_margin_loss(ents, rels, h, r, t, batch_size, margin):
        return torch.stack([O.transe_margin(e, rl, h, r, t, batch_size, margin)[0].reshape(()) for e, rl in zip(ents, rels)])


def _mult_worker(rank, world, port, N, E, R, d, table_bytes):
    """comp_op='mult' (src/jmac_model.py:63-64) destination-sharded: per-edge tables from the gathered source table, whole or per
    block of destinations with recomputation in backward; compared with the single-process oracle in float64."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jmac_b200 import dist as JD
        torch.set_num_threads(1)
        rng = np.random.default_rng(4)
        row, col, typ = rng.integers(0, N, E), rng.integers(0, N, E), rng.integers(0, R, E)
        row[: E // 4] = 2
        ei, et = torch.from_numpy(np.stack([row, col])), torch.from_numpy(typ)
        torch.manual_seed(2)
        x, rel, gout = (torch.randn(N, d, dtype=torch.float64) for N, d in ((N, d), (R, d), (N, d)))
        layers = [make_layer(d, 20), make_layer(d, 21)]
        for l in layers:
            l.comp_op = "mult"
            l.MULT_TABLE_BYTES = table_bytes
        p = [{k: v.detach().clone().requires_grad_(True) for k, v in l.named_parameters()} for l in layers]
        xo, ro = x.clone().requires_grad_(True), rel.clone().requires_grad_(True)
        h = xo
        for pl in p:
            h = O.relation_aware_layer(pl, h, ro, ei, et, slope=0.05, comp_op="mult", bn_state=None, training=True)
        h.backward(gout)
        enc = JD.ShardedEncoder(layers, ei, et, N, R, dist.group.WORLD, ops=CpuOps, chunks=1)
        sh = enc.shard
        relg = rel.clone().requires_grad_(True)
        out = enc.step(x, relg, gout)
        if table_bytes < 1 << 20:
            assert len(layers[0]._mult_blocks(sh.edge_index, sh.edge_type, sh.n_loc, d, CpuOps)) > 1       # the blocked path really ran
        np.testing.assert_allclose(out.detach().numpy(), h.detach().numpy()[sh.lo:sh.hi], atol=1e-10)
        np.testing.assert_allclose(enc._x_loc.grad.numpy(), xo.grad.numpy()[sh.lo:sh.hi], atol=1e-9)
        np.testing.assert_allclose(relg.grad.numpy(), ro.grad.numpy(), atol=1e-9)
        for l, pl in zip(layers, p):
            for k, v in l.named_parameters():
                ref = pl[k].grad if pl[k].grad is not None else torch.zeros_like(pl[k])
                got = v.grad if v.grad is not None else torch.zeros_like(v)
                np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-9, err_msg=k)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("table_bytes", [1 << 30, 40 * 4 * 8])         # whole per-edge table; blocks of <= 40 edges
def test_sharded_mult_encoder_matches_oracle_gloo(table_bytes):
    mp.spawn(_mult_worker, args=(2, _free_port(), 60, 400, 5, 8, table_bytes), nprocs=2, join=True)


def make_layer(d, seed):
    from types import SimpleNamespace
    from jmac_b200.model import RelationAwareLayer
    torch.manual_seed(seed)
    layer = RelationAwareLayer(d, d, rel_dim=d, act=torch.tanh, args=SimpleNamespace(leaky_relu_w=0.05, comp_op="sub")).double()
    with torch.no_grad():
        layer.bn.weight.uniform_(0.5, 1.5)
        layer.bn.bias.normal_(0, 0.1)
    return layer


def _worker(rank, world, port, N, E, R, d, chunks):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jmac_b200 import dist as JD
        torch.set_num_threads(1)
        rng = np.random.default_rng(0)
        row, col, typ = rng.integers(0, N, E), rng.integers(0, N, E), rng.integers(0, R, E)
        row[: E // 4] = 3                                             # a hub: shards become uneven in node count
        ei, et = torch.from_numpy(np.stack([row, col])), torch.from_numpy(typ)
        torch.manual_seed(1)
        x = torch.randn(N, d, dtype=torch.float64)
        rel = torch.randn(R, d, dtype=torch.float64)
        gout = torch.randn(N, d, dtype=torch.float64)
        layers = [make_layer(d, 10), make_layer(d, 11)]
        # ---- single-process oracle (identical on every rank)
        p = [{k: v.detach().clone().requires_grad_(True) for k, v in l.named_parameters()} for l in layers]
        xo, ro = x.clone().requires_grad_(True), rel.clone().requires_grad_(True)
        h = xo
        for pl in p:
            h = O.relation_aware_layer(pl, h, ro, ei, et, slope=0.05, bn_state=None, training=True)
        h.backward(gout)
        # ---- sharded
        enc = JD.ShardedEncoder(layers, ei, et, N, R, dist.group.WORLD, ops=CpuOps, chunks=chunks)
        sh = enc.shard
        assert sh.chunks == chunks and sh.n_max % chunks == 0 and sh.n_total == N
        assert sh.bounds[0] == 0 and sh.bounds[-1] == N and all(b1 > b0 for b0, b1 in zip(sh.bounds, sh.bounds[1:]))
        deg = np.bincount(row, minlength=N)
        assert sh.edge_index.shape[1] == deg[sh.lo:sh.hi].sum()       # every in-edge of an owned node, nothing else
        relg = rel.clone().requires_grad_(True)
        out = enc.step(x, relg, gout)
        np.testing.assert_allclose(out.detach().numpy(), h.detach().numpy()[sh.lo:sh.hi], atol=1e-10)
        np.testing.assert_allclose(enc._x_loc.grad.numpy(), xo.grad.numpy()[sh.lo:sh.hi], atol=1e-9)
        np.testing.assert_allclose(relg.grad.numpy(), ro.grad.numpy(), atol=1e-9)
        for l, pl in zip(layers, p):
            for k, v in l.named_parameters():
                ref = pl[k].grad if pl[k].grad is not None else torch.zeros_like(pl[k])
                got = v.grad if v.grad is not None else torch.zeros_like(v)
                np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-9, err_msg=k)
        # gather / scatter pair on uneven shards
        loc = torch.full((sh.n_loc, 2), float(rank + 1), dtype=torch.float64, requires_grad=True)
        full = JD.GatherRowsFn.apply(loc, sh, dist.group.WORLD)
        assert full.shape[0] == world * sh.n_max
        for g in range(world):
            ng = sh.bounds[g + 1] - sh.bounds[g]
            assert torch.all(full[sh.gathered_row(g, torch.arange(ng))] == g + 1)       # chunk-major layout of the gathered table
        # the local re-layout of a replicated table (first layer: no communication) is the same table
        rep = torch.arange(N, dtype=torch.float64).view(N, 1).repeat(1, 2)
        got = JD.GatherRowsFn.apply(rep[sh.lo:sh.hi].clone(), sh, dist.group.WORLD)
        loc_layout = rep.index_select(0, sh.gather_map("cpu"))
        for g in range(world):
            rows = sh.gathered_row(g, torch.arange(sh.bounds[g + 1] - sh.bounds[g]))
            assert torch.equal(got[rows], loc_layout[rows])
        full.sum().backward()
        assert torch.all(loc.grad == world)                            # every rank contributed a 1 for each of my rows

        # ---- data-parallel completion scorer on row-sharded encoder outputs + a replicated parameter table
        B, K = 10, 3
        hb, rb = rng.integers(0, N, B), rng.integers(0, R, B)
        bh, br = np.tile(hb, K + 1), np.tile(rb, K + 1)
        bt = rng.integers(0, N, (K + 1) * B)
        e0 = x.clone().requires_grad_(True)                            # replicated leaf (the embedding parameter slice)
        e1_full = (x * 0.5 + 1.0).clone().requires_grad_(True)          # stands for the encoder output: every rank owns rows [lo, hi)
        e1_loc = e1_full.detach()[sh.lo:sh.hi].clone().requires_grad_(True)
        rt = [rel.clone().requires_grad_(True), (rel * 2).clone().requires_grad_(True)]
        ref = torch.stack([O.transe_margin(e, rl, T(bh), T(br), T(bt), B, 5.0)[0].reshape(()) for e, rl in zip([e0, e1_full], rt)])
        ref.sum().backward()
        ref_g = [e0.grad.clone(), e1_full.grad.clone(), rt[0].grad.clone(), rt[1].grad.clone()]
        for v in [e0, e1_full] + rt:
            v.grad = None
        e1_g = JD.GatherRowsFn.apply(e1_loc, sh, dist.group.WORLD)
        e1_tab = e1_g.index_select(0, sh.gathered_row(torch.bucketize(torch.arange(N), torch.tensor(sh.bounds)[1:], right=True),
                                                      torch.arange(N) - torch.tensor(sh.bounds)[torch.bucketize(torch.arange(N), torch.tensor(sh.bounds)[1:], right=True)]))
        loss = JD.sharded_transe_margin_loss([e0, e1_tab], rt, bh, br, bt, B, 5.0, dist.group.WORLD, ops=CpuOps)
        np.testing.assert_allclose(loss.detach().numpy(), ref.detach().numpy(), atol=1e-12)     # same value on every rank
        loss.sum().backward()
        JD.all_reduce_grads([e0] + rt, dist.group.WORLD)
        np.testing.assert_allclose(e0.grad.numpy(), ref_g[0].numpy(), atol=1e-12)
        np.testing.assert_allclose(e1_loc.grad.numpy(), ref_g[1].numpy()[sh.lo:sh.hi], atol=1e-12)   # reduce-scattered to the owner
        np.testing.assert_allclose(rt[0].grad.numpy(), ref_g[2].numpy(), atol=1e-12)
        np.testing.assert_allclose(rt[1].grad.numpy(), ref_g[3].numpy(), atol=1e-12)
    finally:
        dist.destroy_process_group()


def T(a):
    return torch.from_numpy(np.asarray(a).astype(np.int64))


@pytest.mark.parametrize("world,chunks", [(2, 1), (3, 1), (2, 3), (3, 2)])
def test_sharded_encoder_matches_oracle_gloo(world, chunks):
    mp.spawn(_worker, args=(world, _free_port(), 61, 400, 5, 16, chunks), nprocs=world, join=True)


def test_positive_share_covers_the_batch():
    from jmac_b200.dist import positive_share
    for B, K1, world in ((10, 4, 3), (7, 2, 8), (1000, 26, 4)):
        M = B * K1
        seen = []
        for r in range(world):
            Bq, b0, b1, sel = positive_share(B, M, r, world)
            assert Bq == B and sel.numel() == (b1 - b0) * K1
            # re-laid-out share keeps the reference pairing: position k * B_r + j is triple k * B + (b0 + j)
            assert torch.equal(sel.view(K1, b1 - b0), torch.arange(K1)[:, None] * B + torch.arange(b0, b1)[None, :])
            seen.append(sel)
        assert torch.equal(torch.sort(torch.cat(seen))[0], torch.arange(M))


def test_partition_balances_in_edges():
    from jmac_b200.dist import partition_by_in_edges, build_shard
    rng = np.random.default_rng(3)
    N, E = 1000, 20000
    row = torch.from_numpy(rng.integers(0, N, E))
    row[:5000] = 17                                                   # hub
    col = torch.from_numpy(rng.integers(0, N, E))
    et = torch.zeros(E, dtype=torch.long)
    for world in (1, 2, 4, 8):
        b = partition_by_in_edges(row, N, world)
        assert len(b) == world + 1 and b[0] == 0 and b[-1] == N
        counts = [int(((row >= b[g]) & (row < b[g + 1])).sum()) for g in range(world)]
        assert sum(counts) == E
        assert max(counts) <= E / world + 5000 + 50                  # the hub cannot be split; everything else balances
        tot = 0
        for g in range(world):
            sh = build_shard(torch.stack([row, col]), et, N, g, world)
            tot += sh.edge_index.shape[1]
            assert int(sh.edge_index[0].max()) < sh.n_loc and int(sh.edge_index[1].max()) < sh.n_src
            # remapped source ids decode back to the original ones
            owner = sh.edge_index[1] // sh.n_max
            back = torch.tensor(sh.bounds)[owner] + sh.edge_index[1] % sh.n_max
            mine = (row >= sh.lo) & (row < sh.hi)
            assert torch.equal(back, col[mine])
            # chunked gathered layout: the replicated table re-laid-out through gather_map holds the right row under every edge
            sh3 = build_shard(torch.stack([row, col]), et, N, g, world, chunks=3)
            assert sh3.n_max % 3 == 0 and torch.equal(sh3.gather_map("cpu")[sh3.edge_index[1]], col[mine])
        assert tot == E
