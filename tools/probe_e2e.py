"""Breakdown of the e2e step (H2D, plan build, compute) -- diagnostic, not a bench line."""
import sys, time, torch
sys.path.insert(0, ".")
from types import SimpleNamespace
from jmac_b200 import RelationAwareLayer, functional as JF
dev = torch.device("cuda")
N, R, E, D = 2_000_000, 2000, 40_000_000, 256
g = torch.Generator(device=dev).manual_seed(0)
ei = torch.randint(0, N, (2, E), device=dev, generator=g); et = torch.randint(0, R, (E,), device=dev, generator=g)
x0 = torch.randn(N, D, device=dev, generator=g) * 0.1; rel = torch.randn(R, D, device=dev, generator=g) * 0.1
gout = torch.randn(N, D, device=dev, generator=g)
layers = [RelationAwareLayer(D, D, rel_dim=D, act=torch.tanh, args=SimpleNamespace(leaky_relu_w=0.05, comp_op="sub")).to(dev) for _ in range(2)]
h_x, h_r, h_ei, h_et = x0.cpu().pin_memory(), rel.cpu().pin_memory(), ei.cpu().pin_memory(), et.cpu().pin_memory()
del ei, et, x0
def T(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
ms, (dx, dr, dei, det) = T(lambda: (h_x.to(dev, non_blocking=True), h_r.to(dev, non_blocking=True), h_ei.to(dev, non_blocking=True), h_et.to(dev, non_blocking=True)))
print(f"H2D 3.0 GB: {ms:.1f} ms")
ms, _ = T(lambda: h_ei.to(torch.int32))
print(f"host int64->int32 cast of edge_index: {ms:.1f} ms")
ms, plan = T(lambda: JF.GraphPlan(dei, det, N, R + 1))
print(f"plan build: {ms:.1f} ms")
def comp():
    JF.clear_plan_cache()
    x = dx.detach().requires_grad_(True); r = dr.detach().requires_grad_(True)
    h = layers[1](layers[0](x, r, dei, det), r, dei, det)
    loss = (h * gout).sum(); loss.backward(); return float(loss.item())
ms, _ = T(comp)
print(f"compute incl. 1 plan build via cache: {ms:.1f} ms")
