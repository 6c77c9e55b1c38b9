"""Accuracy of the 3xTF32 node GEMM variants against float64, and what it means for the LeakyReLU sign decisions
of the attention logit z = P[i] + Q[j] - S[t] ("kinks").

profiles/gemm_accuracy_r02.jsonl was recorded at commit a443bbe, where the kernel still had an experiment switch
(JMAC_GEMM_MODE, removed afterwards because it changed nothing):
mode 0: lo = x - hi fed to the tensor core as is (its low 13 mantissa bits are dropped by the MMA), 3 products  [what ships]
mode 1: lo rounded to TF32 (RN) before the MMA, 3 products
mode 3: lo rounded to TF32 + the lo.lo product (4 MMAs)
All three: rms error 1.8e-7 of sum|a||b| (cuBLAS fp32: 2.8e-8), max 1.1-1.2e-6 -- the error is the tensor core's accumulation,
not the operand split.
Prints one JSON line: max / rms error of P = X W in units of sum|a||b| and of an fp32 FMA chain (cuBLAS fp32,
TF32 disabled), and the number of sign decisions of z that differ from float64 on a 5000-node / 30001-edge graph."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmac_b200 import functional as JF  # noqa: E402


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    M, K, N = 8192, 256, 256
    X = torch.randn(M, K, device="cuda", generator=g) * 0.5
    W = torch.empty(K, N, device="cuda")
    torch.nn.init.xavier_normal_(W)
    ref = X.double() @ W.double()
    bound = X.abs().double() @ W.abs().double()
    ours = JF.gemm_nn(X, W).double()
    lib = (X @ W).double()
    out = {"mode": int(os.environ.get("JMAC_GEMM_MODE", "0")),
           "max_err_over_bound": float(((ours - ref).abs() / bound).max()), "rms_err_over_bound": float((((ours - ref) / bound) ** 2).mean().sqrt()),
           "cublas_fp32_max_err_over_bound": float(((lib - ref).abs() / bound).max()),
           "cublas_fp32_rms_err_over_bound": float((((lib - ref) / bound) ** 2).mean().sqrt()),
           "mean_signed_err_over_bound": float(((ours - ref) / bound).mean())}
    # sign decisions of z on a graph of the size the budget test uses (tests/test_gpu_layer.py, N=5000, E=30001)
    n, e, r = 5000, 30001, 24
    Xn = torch.randn(n, K, device="cuda", generator=g) * 0.5
    Rp = torch.randn(r, K, device="cuda", generator=g) * 0.5
    Wa = torch.empty(2 * K, N, device="cuda")
    torch.nn.init.xavier_normal_(Wa)
    i = torch.randint(0, n, (e,), device="cuda", generator=g)
    j = torch.randint(0, n, (e,), device="cuda", generator=g)
    t = torch.randint(0, r, (e,), device="cuda", generator=g)
    z64 = torch.cat([Xn[i], Xn[j] - Rp[t]], 1).double() @ Wa.double()
    z_ref32 = torch.cat([Xn[i], Xn[j] - Rp[t]], 1) @ Wa                                   # the reference's fp32 formulation
    Wt, Wb = Wa[:K], Wa[K:]
    P = JF.gemm_nn(Xn, Wt)
    Q = JF.gemm_nn(Xn, Wb)
    S = Rp @ Wb
    z_ours = P[i] + Q[j] - S[t]
    out["sign_flips_vs_f64"] = {"ours": int(((z_ours > 0) != (z64 > 0)).sum()), "reference_fp32": int(((z_ref32 > 0) != (z64 > 0)).sum()),
                                "decisions": int(z64.numel())}
    out["z_rms_err"] = {"ours": float((z_ours.double() - z64).pow(2).mean().sqrt()), "reference_fp32": float((z_ref32.double() - z64).pow(2).mean().sqrt())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
