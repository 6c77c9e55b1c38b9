import sys, torch
sys.path.insert(0, ".")
from jmac_b200 import functional as JF, _lib
lib = _lib.load()
torch.manual_seed(0)
for M in (16, 48, 5000):
    X = torch.randn(M, 256, device="cuda"); G = torch.randn(M, 256, device="cuda")
    D = torch.full((256, 256), 7.0, device="cuda")
    ws = torch.full((148 * 256 * 256,), 3.0, device="cuda")
    rc = lib.jmac_gemm_tn_tf32x3(_lib.ptr(X), 256, _lib.ptr(G), 256, _lib.ptr(D), M, 256, 256, _lib.ptr(ws), _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = X.t() @ G
    print("M", M, "rc", rc, "D[0,:4]", D[0, :4].tolist(), "ref", ref[0, :4].tolist())
    print("   ws first", ws[:4].tolist(), "ws nonzero frac", float((ws[:256*256] != 0).float().mean()), "maxerr", float((D - ref).abs().max()))
