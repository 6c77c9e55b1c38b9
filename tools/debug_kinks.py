"""How many LeakyReLU sign decisions differ between the 3xTF32 node products and an fp64 product (large hub case)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jmac_b200 import functional as JF

N, E, R, d, hub = 5000, 30001, 40, 256, 2000
rng = np.random.default_rng(1234 + N)
row = rng.integers(0, N, E); col = rng.integers(0, N, E); typ = rng.integers(0, R, E)
row[:hub] = 5; col[hub:2 * hub] = 9; typ[:hub // 2] = 1
row[row == 0] = 1; row[row == N - 1] = 2
torch.manual_seed(1234 + N)
x = (torch.randn(N, d) * 0.5).cuda(); rp = (torch.randn(R + 1, d) * 0.5).cuda()
torch.manual_seed(7)
W = torch.empty(2 * d, d); torch.nn.init.xavier_normal_(W); W = W.cuda()
i, j, t = (torch.from_numpy(a).cuda() for a in (row, col, typ))
for name, mm in (("3xTF32", lambda a, b: JF.gemm_nt(a, b.t().contiguous())), ("torch.mm fp32", torch.mm)):
    P, Q, S = mm(x, W[:d]), mm(x, W[d:]), rp @ W[d:]
    z = P[i] + Q[j] - S[t]
    z64 = (x.double() @ W[:d].double())[i] + (x.double() @ W[d:].double())[j] - (rp.double() @ W[d:].double())[t]
    flips = (z > 0) != (z64 > 0)
    print(f"{name}: {int(flips.sum())} of {z.numel()} sign decisions differ from fp64; largest |z64| among them {float(z64[flips].abs().max()) if flips.any() else 0:.2e}; "
          f"max |z - z64| {float((z.double() - z64).abs().max()):.2e}; edges into the hub among them {int(flips[:hub].any(1).sum())}")
