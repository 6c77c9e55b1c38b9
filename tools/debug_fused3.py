"""Check every gemm_nt / gemm_tn call of one layer step (large hub case) against an fp64 product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from jmac_b200 import RelationAwareLayer, functional as JF

N, E, R, d, hub = 5000, 30001, 40, 256, 2000
rng = np.random.default_rng(1234 + N)
row = rng.integers(0, N, E); col = rng.integers(0, N, E); typ = rng.integers(0, R, E)
row[:hub] = 5; col[hub:2 * hub] = 9; typ[:hub // 2] = 1
row[row == 0] = 1; row[row == N - 1] = 2
torch.manual_seed(1234 + N)
x0 = torch.randn(N, d) * 0.5; r0 = torch.randn(R, d) * 0.5; gout = torch.randn(N, d)
ei = np.stack([row, col]).astype(np.int64)
orig_nt, orig_tn = JF.gemm_nt, JF.gemm_tn
calls = []

def nt(A, B, add1=None, add2=None, out=None):
    ref = A.double() @ B.double().t()
    if add1 is not None: ref = ref + add1.double()
    if add2 is not None: ref = ref + add2.double()
    bound = A.double().abs() @ B.double().abs().t() + (add1.double().abs() if add1 is not None else 0) + 1e-30
    C = orig_nt(A, B, add1, add2, out)
    err = (C.double() - ref).abs()
    i = int(err.argmax()) // C.shape[1]
    print(f"gemm_nt M={A.shape[0]} K={A.shape[1]} N={B.shape[0]} add={add1 is not None},{add2 is not None}: max abs err {float(err.max()):.3e} "
          f"max err/bound {float((err / bound).max()):.3e} worst row {i} |A row|max {float(A[i].abs().max()):.3e} |A|max {float(A.abs().max()):.3e} "
          f"A row min nonzero {float(A[i][A[i] != 0].abs().min()) if (A[i] != 0).any() else 0:.3e}")
    return C

JF.gemm_nt = nt
torch.manual_seed(7)
layer = RelationAwareLayer(d, d, rel_dim=d, act=torch.tanh, args=SimpleNamespace(leaky_relu_w=0.05, comp_op="sub")).cuda()
x = x0.cuda().requires_grad_(True); r = r0.cuda().requires_grad_(True)
out = layer(x, r, torch.from_numpy(ei).cuda(), torch.from_numpy(typ).cuda())
(out * gout.cuda()).sum().backward()
