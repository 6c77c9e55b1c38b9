"""Which resource bounds edge_fwd? (diagnostic) baseline / single relation (no L2 relation traffic) / interleaved Q|X rows."""
import sys, torch
sys.path.insert(0, ".")
from jmac_b200 import functional as JF, _lib
lib = _lib.load(); ptr = _lib.ptr; st = _lib.stream_ptr
dev = torch.device("cuda"); N, R, E, d = 2_000_000, 2000, 40_000_000, 256
g = torch.Generator(device=dev).manual_seed(0)
ei = torch.randint(0, N, (2, E), device=dev, generator=g); et = torch.randint(0, R, (E,), device=dev, generator=g)
P = torch.randn(N, d, device=dev); S = torch.randn(R + 1, d, device=dev); Rp = torch.randn(R + 1, d, device=dev); a = torch.randn(d, device=dev)
T = torch.randn(2 * N + 1, d, device=dev)
agg = torch.empty(N, d, device=dev); lse = torch.empty(N, device=dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
def run(plan, Q, X):
    return t(lambda: _lib.check(lib.jmac_edge_fwd(plan.handle, ptr(P), ptr(Q), ptr(X), ptr(S), ptr(Rp), ptr(a), 0.05, d, ptr(agg), ptr(lse), st())))
Q, X = T[:N].contiguous(), T[N:2 * N].contiguous()
plan = JF.GraphPlan(ei, et, N, R + 1)
print("baseline            %.2f ms" % run(plan, Q, X))
plan0 = JF.GraphPlan(ei, torch.zeros_like(et), N, R + 1)
print("single relation     %.2f ms" % run(plan0, Q, X))
ei2 = torch.stack([ei[0], ei[1] * 2])
plan2 = JF.GraphPlan(ei2, et, N, R + 1, n_src=2 * N)
print("interleaved Q|X     %.2f ms" % run(plan2, T[:2 * N], T[1:]))
plan3 = JF.GraphPlan(ei2, torch.zeros_like(et), N, R + 1, n_src=2 * N)
print("interleaved + 1 rel %.2f ms" % run(plan3, T[:2 * N], T[1:]))
# sorted sources inside each destination segment do not change anything for uniform graphs; skip
