"""Scorer fwd+bwd time at the bench's size (CUDA events)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmac_b200 import functional as JF
N, R, D, B, K = 2_000_000, 2_000, 256, 65_536, 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
e = [(torch.randn(N, D, device=dev, generator=g) * 0.1).requires_grad_(True) for _ in range(2)]
r = [(torch.randn(R, D, device=dev, generator=g) * 0.1).requires_grad_(True) for _ in range(2)]
rng = np.random.default_rng(0)
hh = torch.from_numpy(np.tile(rng.integers(0, N, B), K + 1)).to(dev)
rr = torch.from_numpy(np.tile(rng.integers(0, R, B), K + 1)).to(dev)
tt = torch.from_numpy(rng.integers(0, N, (K + 1) * B)).to(dev)
def step(bwd=True):
    for t_ in e + r:
        t_.grad = None
    l = JF.transe_margin_loss(e, r, hh, rr, tt, B, 5.0)
    if bwd:
        l.sum().backward()
for bwd in (True, False):
    for _ in range(3):
        step(bwd)
    torch.cuda.synchronize()
    s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        step(bwd)
    f.record()
    torch.cuda.synchronize()
    print(f"variant {os.environ.get('JMAC_SCORER_VARIANT', '0')} {'fwd+bwd' if bwd else 'fwd(record)'}: {s.elapsed_time(f) / 10:.3f} ms")
