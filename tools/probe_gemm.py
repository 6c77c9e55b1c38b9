import sys, torch
sys.path.insert(0, ".")
from jmac_b200 import functional as JF
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
M, K, N = 2_000_000, 256, 256
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") * 0.1; G = torch.randn(M, N, device="cuda")
print("NT tf32x3 kernel  %.3f ms" % t(lambda: JF.gemm_nt(A, B)))
print("NT torch fp32 mm  %.3f ms" % t(lambda: A @ B.t()))
print("TN tf32x3 kernel  %.3f ms" % t(lambda: JF.gemm_tn(A, G)))
print("TN torch fp32 mm  %.3f ms" % t(lambda: A.t() @ G))
