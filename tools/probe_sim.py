"""Quick perf probe of the similarity kernel (not a bench line)."""
import sys, time, torch
sys.path.insert(0, ".")
from jmac_b200 import finding as F
def run(n1, n2, d=256, k=10):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.nn.functional.normalize(torch.randn(n1, d, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
    B = torch.nn.functional.normalize(torch.randn(n2, d, device="cuda", generator=g)).to(torch.bfloat16).contiguous()
    r1 = torch.rand(n1, device="cuda"); r2 = torch.rand(n2, device="cuda"); gv = torch.zeros(n1, device="cuda")
    for name, fn in (("topk", lambda: F.sim_topk(A, B, k)), ("count_only", lambda: F.sim_topk(A, B, 0, 2.0, r1, r2, gold_val=gv, want_count=True)),
                     ("csls_top1_count", lambda: F.sim_topk(A, B, 1, 2.0, r1, r2, gold_val=gv, want_count=True)),
                     ("plain_nothing", lambda: F.sim_topk(A, B, 0, want_count=False))):
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        print(f"n1={n1} n2={n2} {name}: {ms:.2f} ms  {2*n1*n2*d/ms/1e9:.1f} TFLOP/s", flush=True)
for n in (15000, 131072, 524288):
    run(n, n)
run(1048576, 1048576)
