"""Fused attention-input node vs the composed path (node_mm + edge_phase) on the same inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from jmac_b200 import RelationAwareLayer, functional as JF

def run(N, E, R, d, hub, fused):
    rng = np.random.default_rng(1234 + N)
    row = rng.integers(0, N, E); col = rng.integers(0, N, E); typ = rng.integers(0, R, E)
    if hub:
        row[:hub] = 5; col[hub:2 * hub] = 9; typ[:hub // 2] = 1
    torch.manual_seed(1234 + N)
    x = (torch.randn(N, d) * 0.5).cuda().requires_grad_(True)
    r = (torch.randn(R, d) * 0.5).cuda().requires_grad_(True)
    gout = torch.randn(N, d).cuda()
    torch.manual_seed(7)
    args = SimpleNamespace(leaky_relu_w=0.05, comp_op="sub")
    layer = RelationAwareLayer(d, d, rel_dim=d, act=torch.tanh, args=args).cuda()
    if not fused:
        JF.attn_edge_eligible = lambda X, d: False
    out = layer(x, r, torch.from_numpy(np.stack([row, col])).cuda(), torch.from_numpy(typ).cuda())
    (out * gout).sum().backward()
    return out.detach(), x.grad, r.grad, {k: p.grad for k, p in layer.named_parameters()}

for shape in ((5000, 30001, 40, 256, 2000), (5000, 30000, 40, 256, 0), (4200, 9000, 7, 128, 0), (8192, 40000, 5, 256, 0)):
    import importlib
    importlib.reload(JF)
    import jmac_b200.model as M
    M.JF = JF
    a = run(*shape, fused=True)
    importlib.reload(JF); M.JF = JF
    b = run(*shape, fused=False)
    rel = lambda u, v: float((u - v).abs().max() / v.abs().max().clamp_min(1e-30))
    print(shape, "out", rel(a[0], b[0]), "dX", rel(a[1], b[1]), "dR", rel(a[2], b[2]), {k: round(rel(a[3][k], b[3][k]), 7) for k in a[3] if a[3][k] is not None})
    bad = ((a[1] - b[1]).abs() > 1e-4 * b[1].abs().max()).nonzero()
    if len(bad):
        print("   bad rows", bad[:, 0].unique()[:20].tolist(), "cols", bad[:, 1].unique()[:20].tolist(), "count", len(bad))
