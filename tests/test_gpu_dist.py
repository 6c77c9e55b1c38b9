"""Multi-GPU (NCCL) parity: the destination-sharded encoder and the row-sharded alignment evaluation on 2 GPUs against the single-GPU CUDA
path on the same inputs.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a 1-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_layers(d):
    from types import SimpleNamespace
    from jmac_b200 import RelationAwareLayer
    torch.manual_seed(5)
    return [RelationAwareLayer(d, d, rel_dim=d, act=torch.tanh, args=SimpleNamespace(leaky_relu_w=0.05, comp_op="sub")) for _ in range(2)]


def _worker(rank, world, port, N, E, R, d):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from jmac_b200 import dist as JD
        rng = np.random.default_rng(0)
        row, col, typ = rng.integers(0, N, E), rng.integers(0, N, E), rng.integers(0, R, E)
        row[: E // 5] = 11
        ei, et = torch.from_numpy(np.stack([row, col])).to(dev), torch.from_numpy(typ).to(dev)
        torch.manual_seed(1)
        x, rel, gout = torch.randn(N, d) * 0.5, torch.randn(R, d) * 0.5, torch.randn(N, d)
        # single-GPU reference (each rank computes it on its own device)
        ref_layers = [l.to(dev) for l in _make_layers(d)]
        xs, rs = x.to(dev).requires_grad_(True), rel.to(dev).requires_grad_(True)
        h = ref_layers[1](ref_layers[0](xs, rs, ei, et), rs, ei, et)
        h.backward(gout.to(dev))
        # sharded
        layers = [l.to(dev) for l in _make_layers(d)]
        enc = JD.ShardedEncoder(layers, ei, et, N, R, dist.group.WORLD)
        rg = rel.to(dev).requires_grad_(True)
        out = enc.step(x.to(dev), rg, gout.to(dev))
        sh = enc.shard
        torch.testing.assert_close(out, h.detach()[sh.lo:sh.hi], atol=1e-4, rtol=1e-4)   # two stacked layers; the hub row (sqrt(deg) = 110) amplifies fp32 path differences

        def close(a, b, name):
            err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
            assert err < 1e-2, f"{name}: rel err {err:.2e}"      # fp32-vs-fp32: each side carries ~1e-3 softmax-backward cancellation noise

        close(enc._x_loc.grad, xs.grad[sh.lo:sh.hi], "dX")
        close(rg.grad, rs.grad, "dR")
        for l, lr in zip(layers, ref_layers):
            for (k, v), (_, vr) in zip(l.named_parameters(), lr.named_parameters()):
                if k == "loop_rel":
                    continue                                       # mathematically zero (fp32 noise only)
                close(v.grad, vr.grad, k)
            torch.testing.assert_close(l.bn.running_mean, lr.bn.running_mean, atol=1e-6, rtol=1e-5)
            torch.testing.assert_close(l.bn.running_var, lr.bn.running_var, atol=1e-6, rtol=1e-5)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("N,E", [(5000, 60000),          # shards below 4096 rows: composed path (library GEMMs for the local rows)
                                 (12000, 100001)])       # shards of ~6000 rows: the fused attention-input node on a gathered source table
def test_sharded_encoder_matches_single_gpu_nccl(N, E):
    mp.spawn(_worker, args=(2, _free_port(), N, E, 30, 256), nprocs=2, join=True)


def _align_worker(rank, world, port, n1, n2, d, csls):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from jmac_b200 import finding as F
        g = torch.Generator(device=dev).manual_seed(3)
        n = min(n1, n2)
        e1 = torch.nn.functional.normalize(torch.randn(n1, d, device=dev, generator=g))
        e2 = torch.nn.functional.normalize(torch.randn(n2, d, device=dev, generator=g))
        e2[:n] = torch.nn.functional.normalize(e1[:n] + 1.5 * torch.nn.functional.normalize(torch.randn(n, d, device=dev, generator=g)))
        A, B = F.to_bf16_rows(e1[:n]), F.to_bf16_rows(e2)            # gold column of row i = i
        ranks1, top1 = F.alignment_ranks_bf16(A, B, csls_k=csls)                       # this rank alone
        ranks2, top2 = F.alignment_ranks_bf16(A, B, csls_k=csls, group=dist.group.WORLD)  # rows sharded over the ranks
        assert torch.equal(ranks1, ranks2) and torch.equal(top1, top2)                  # same kernels on disjoint rows: bit-identical
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("n1,n2,csls", [(3000, 3000, 10), (1025, 2500, 0)])
def test_row_sharded_alignment_matches_single_gpu_nccl(n1, n2, csls):
    mp.spawn(_align_worker, args=(2, _free_port(), n1, n2, 256, csls), nprocs=2, join=True)
