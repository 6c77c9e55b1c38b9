"""Where does the dX error of the large hub case come from?  GPU paths vs the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from jmac_b200 import RelationAwareLayer, functional as JF
import jmac_b200.model as M
from oracle import jmac_oracle as O

N, E, R, d, hub = 5000, 30001, 40, 256, 2000
rng = np.random.default_rng(1234 + N)
row = rng.integers(0, N, E); col = rng.integers(0, N, E); typ = rng.integers(0, R, E)
row[:hub] = 5; col[hub:2 * hub] = 9; typ[:hub // 2] = 1
row[row == 0] = 1; row[row == N - 1] = 2
torch.manual_seed(1234 + N)
x0 = torch.randn(N, d) * 0.5; r0 = torch.randn(R, d) * 0.5; gout = torch.randn(N, d)
ei = np.stack([row, col]).astype(np.int64)

def gpu(mode):
    torch.manual_seed(7)
    layer = RelationAwareLayer(d, d, rel_dim=d, act=torch.tanh, args=SimpleNamespace(leaky_relu_w=0.05, comp_op="sub")).cuda()
    tc, tn, ae = JF._tc_eligible, JF._tn_eligible, JF.attn_edge_eligible
    if mode != "fused":
        JF.attn_edge_eligible = lambda X, d: False
    if mode == "torchmm":
        JF._tc_eligible = lambda *a: False
        JF._tn_eligible = lambda *a: False
    x = x0.cuda().requires_grad_(True); r = r0.cuda().requires_grad_(True)
    out = layer(x, r, torch.from_numpy(ei).cuda(), torch.from_numpy(typ).cuda())
    (out * gout.cuda()).sum().backward()
    JF._tc_eligible, JF._tn_eligible, JF.attn_edge_eligible = tc, tn, ae
    return layer, out.detach().cpu(), x.grad.cpu()

layer, o_f, gx_f = gpu("fused")
_, o_c, gx_c = gpu("composed")
_, o_t, gx_t = gpu("torchmm")
p = {k: v.detach().cpu().double().clone().requires_grad_(True) for k, v in layer.named_parameters()}
st = O.BNState(torch.zeros(d), torch.ones(d))
xo = x0.double().requires_grad_(True); ro = r0.double().requires_grad_(True)
out = O.relation_aware_layer(p, xo, ro, torch.from_numpy(ei), torch.from_numpy(typ), slope=0.05, bn_state=st, training=True)
(out * gout.double()).sum().backward()
g64 = xo.grad
mx = float(g64.abs().max())
for name, g in (("fused", gx_f), ("composed+gemm", gx_c), ("composed+torch.mm", gx_t)):
    err = (g.double() - g64).abs()
    i = int(err.argmax()) // d
    print(name, "max abs err", float(err.max()), "rel to max", float(err.max()) / mx, "at row", i, "row max |g|", float(g64[i].abs().max()),
          "rows with err > 1e-5*max:", int((err.max(1)[0] > 1e-5 * mx).sum()))
print("max |g64|", mx, "at row", int(g64.abs().max(1)[0].argmax()), "out err fused", float((o_f.double() - out.detach()).abs().max()),
      "torchmm", float((o_t.double() - out.detach()).abs().max()))
