"""Split (stream-K) vs unsplit launch of the bf16 similarity kernel at the per-rank shape of the 8-GPU C5 run (131072 x 1M)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmac_b200 import finding as F
g = torch.Generator(device="cuda").manual_seed(0)
A = F.to_bf16_rows(torch.nn.functional.normalize(torch.randn(131072, 256, device="cuda", generator=g)))
B = F.to_bf16_rows(torch.nn.functional.normalize(torch.randn(1048576, 256, device="cuda", generator=g)))
for split in (True, False, True, False):
    F.sim_topk(A, B, 10, split=split)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        F.sim_topk(A, B, 10, split=split)
    e.record()
    torch.cuda.synchronize()
    print(f"split={split}: {s.elapsed_time(e) / 3:.2f} ms per pass")
