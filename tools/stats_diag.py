"""Which of the split / unsplit softmax-statistics sweeps is off when they disagree (float64 truth from the same operands)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmac_b200 import finding as F

torch.manual_seed(0)
for n1, n2, d in ((300, 5000, 64), (4500, 30000, 256)):
    e1 = torch.nn.functional.normalize(torch.randn(n1, d, device="cuda"))
    e2 = torch.nn.functional.normalize(torch.randn(n2, d, device="cuda"))
    m = min(n1, n2)
    e2[:m] = torch.nn.functional.normalize(e1[:m] + torch.nn.functional.normalize(torch.randn(m, d, device="cuda")))
    for prec in ("fp32", "bf16"):
        A, B = F.to_rows(e1, precision=prec), F.to_rows(e2, precision=prec)
        a = (A.hi.double() + A.lo.double()) if A.f32 else A.hi.double()
        b = (B.hi.double() + B.lo.double()) if B.f32 else B.hi.double()
        mask1 = (torch.rand(n1, device="cuda") < 0.2)
        mask2 = (torch.rand(n2, device="cuda") < 0.2)
        X = 20.0 * (a @ b.t())
        X[mask1] = -20.0
        X[:, mask2] = -20.0
        M = X.max(1)[0]
        Z = torch.exp(X - M[:, None]).sum(1)
        W = (torch.exp(X - M[:, None]) * (X - M[:, None])).sum(1)
        for split in (True, False):
            s = F.softmax_stats(A, B, 20.0, row_mask=mask1.float(), col_mask=mask2.float(), split=split)
            ez = ((s["Z"].double() - Z).abs() / Z)
            ew = ((s["W"].double() - W).abs() / W.abs().clamp_min(1e-3))
            i = int(ez.argmax())
            print(f"{n1}x{n2} d={d} {prec} split={split}: Z rel err max {float(ez.max()):.2e} mean {float(ez.mean()):.2e} | W max {float(ew.max()):.2e}"
                  f" | worst row {i}: Z={float(s['Z'][i]):.6f} truth {float(Z[i]):.6f} max {float(s['max'][i]):.4f} truth {float(M[i]):.4f} masked_row={bool(mask1[i])}")
