import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from types import SimpleNamespace
from jmac_b200 import RelationAwareLayer, dist as JD
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
N, E, R, d = 5000, 60000, 30, 256
rng = np.random.default_rng(0)
row, col, typ = rng.integers(0, N, E), rng.integers(0, N, E), rng.integers(0, R, E); row[:E // 5] = 11
ei, et = torch.from_numpy(np.stack([row, col])).to(dev), torch.from_numpy(typ).to(dev)
torch.manual_seed(1); x, rel = torch.randn(N, d) * 0.5, torch.randn(R, d) * 0.5
torch.manual_seed(5); layers = [RelationAwareLayer(d, d, rel_dim=d, act=torch.tanh, args=SimpleNamespace(leaky_relu_w=0.05, comp_op="sub")).to(dev) for _ in range(2)]
torch.manual_seed(5); ref = [RelationAwareLayer(d, d, rel_dim=d, act=torch.tanh, args=SimpleNamespace(leaky_relu_w=0.05, comp_op="sub")).to(dev) for _ in range(2)]
xs, rs = x.to(dev), rel.to(dev)
with torch.no_grad():
    h1 = ref[0](xs, rs, ei, et); h2 = ref[1](h1, rs, ei, et)
    enc = JD.ShardedEncoder(layers, ei, et, N, R, dist.group.WORLD); sh = enc.shard
    xl = xs[sh.lo:sh.hi].contiguous()
    o1 = JD.sharded_layer_forward(layers[0], xl, rs, sh, enc.plan, enc.group, enc.ops, xs)
    o1b = JD.sharded_layer_forward(layers[0], xl, rs, sh, enc.plan, enc.group, enc.ops, None)
    o2 = JD.sharded_layer_forward(layers[1], o1b, rs, sh, enc.plan, enc.group, enc.ops, None)
    print(rank, "bounds", sh.bounds, "n_max", sh.n_max, "L1(local gather) err", float((o1 - h1[sh.lo:sh.hi]).abs().max()),
          "L1(allgather) err", float((o1b - h1[sh.lo:sh.hi]).abs().max()), "L2 err", float((o2 - h2[sh.lo:sh.hi]).abs().max()), flush=True)
dist.destroy_process_group()
