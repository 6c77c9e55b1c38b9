"""GPU parity of the 3xTF32 tcgen05 node GEMM against float64 matmul (the bar is fp32 accuracy:
a plain TF32 product has ~5e-4 relative error and would miss north_star's 1e-4 tolerance)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("M,K,N", [(5000, 256, 256), (4097, 64, 128), (300, 256, 32), (70000, 256, 256), (1000, 16, 96), (257, 128, 224),
                                   (14000, 256, 256), (13176, 256, 192),       # DBP-5L node counts: row blocks shared by 2 CTAs (N / 2 columns each)
                                   (5000, 320, 320), (70000, 320, 320), (4100, 304, 512), (300, 48, 288)])   # N > 256: always column parts (d = 300 padded to 320)
def test_gemm_tf32x3_matches_fp64(M, K, N):
    from jmac_b200 import functional as JF
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device="cuda") * torch.exp(torch.randn(M, 1, device="cuda"))       # rows of very different scale
    B = torch.randn(N, K, device="cuda") * 0.3
    C = JF.gemm_nt(A, B)
    ref = A.double() @ B.double().t()
    err = float(((C.double() - ref).abs() / (A.double().abs() @ B.double().abs().t() + 1e-30)).max())    # error vs the magnitude sum
    err32 = float((((A @ B.t()).double() - ref).abs() / (A.double().abs() @ B.double().abs().t() + 1e-30)).max())
    assert err < 2e-6, f"3xTF32 error {err:.2e} (library fp32: {err32:.2e})"
    assert relerr(C, ref) < 5e-6


@pytest.mark.parametrize("M,KO,N", [(5000, 256, 256), (4099, 128, 96), (70001, 256, 32), (16, 256, 256), (40000, 128, 256),
                                    (5000, 64, 160), (4099, 224, 32), (5000, 320, 320), (20000, 512, 288)])   # partial 128-row halves; column blocks
def test_gemm_tn_tf32x3_matches_fp64(M, KO, N):
    """Weight-gradient product X^T G (MN-major operands, split-M partials, deterministic)."""
    from jmac_b200 import functional as JF
    torch.manual_seed(M + KO)
    X = torch.randn(M, KO, device="cuda") * torch.exp(0.5 * torch.randn(M, 1, device="cuda"))
    G = torch.randn(M, N, device="cuda")
    D = JF.gemm_tn(X, G)
    ref = X.double().t() @ G.double()
    scale = X.double().abs().t() @ G.double().abs()
    assert float(((D.double() - ref).abs() / scale).max()) < 2e-6
    assert torch.equal(D, JF.gemm_tn(X, G))                      # fixed-order partial sums: bit-stable


def test_node_mm_autograd_matches_torch():
    from jmac_b200 import functional as JF
    torch.manual_seed(0)
    X = torch.randn(9000, 256, device="cuda", requires_grad=True)
    W = (torch.randn(256, 256, device="cuda") * 0.1).requires_grad_(True)
    G = torch.randn(9000, 256, device="cuda")
    Y = JF.node_mm(X, W)
    Y.backward(G)
    Xr, Wr = X.detach().double().requires_grad_(True), W.detach().double().requires_grad_(True)
    (Xr @ Wr).backward(G.double())
    assert relerr(Y.detach(), (Xr @ Wr).detach()) < 5e-6
    assert relerr(X.grad, Xr.grad) < 5e-6
    assert relerr(W.grad, Wr.grad) < 1e-5
    # small / odd shapes fall back to the library product (same semantics)
    x2 = torch.randn(100, 300, device="cuda")
    w2 = torch.randn(300, 300, device="cuda")
    assert torch.equal(JF.node_mm(x2, w2), torch.mm(x2, w2))


@pytest.mark.parametrize("M,K,N", [(5000, 256, 256), (4321, 128, 96), (513, 64, 32), (14000, 256, 256)])
def test_gemm_epilogue_addends_and_in_place_accumulate(M, K, N):
    """jmac_gemm_tf32x3_add: C = A B^T + D1 + D2 with 256-bit epilogue accesses; D1 may alias C (gradient accumulation)."""
    from jmac_b200 import functional as JF
    torch.manual_seed(M)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") * 0.3
    D1, D2 = torch.randn(M, N, device="cuda"), torch.randn(M, N, device="cuda") * 3
    ref = A.double() @ B.double().t() + D1.double() + D2.double()
    bound = A.double().abs() @ B.double().abs().t() + D1.double().abs() + D2.double().abs()
    C = JF.gemm_nt(A, B, add1=D1, add2=D2)
    assert float(((C.double() - ref).abs() / bound).max()) < 2e-6
    acc = D1.clone()
    out = JF.gemm_nt(A, B, add1=acc, add2=D2, out=acc)                 # in place over the first addend
    assert out.data_ptr() == acc.data_ptr() and torch.equal(acc, C)
    only2 = JF.gemm_nt(A, B, add2=D2)                                   # a single addend given as the second
    assert float(((only2.double() - (ref - D1.double())).abs() / bound).max()) < 2e-6
    with pytest.raises(ValueError):
        JF.gemm_nt(A, B, add1=D1[:, : N // 2])


@pytest.mark.parametrize("d", [32, 256, 300, 512])
def test_row_normalize_matches_torch(d):
    """F.normalize(x, 2, -1) forward and backward in one kernel each, including a zero row (the eps clamp: no norm gradient)."""
    from jmac_b200 import functional as JF
    torch.manual_seed(d)
    x = torch.randn(777, d, device="cuda")
    x[5] = 0
    x[6] *= 1e-14                                           # below eps = 1e-12: y = x / eps
    g = torch.randn(777, d, device="cuda")
    a = x.clone().requires_grad_(True)
    b = x.double().clone().requires_grad_(True)
    ya = JF.row_normalize(a)
    yb = torch.nn.functional.normalize(b, 2, -1)
    (ya * g).sum().backward()
    (yb * g.double()).sum().backward()
    torch.testing.assert_close(ya.detach().double(), yb.detach(), rtol=1e-6, atol=1e-7)
    scale = float(b.grad.abs().max())
    rows = torch.ones(777, dtype=torch.bool, device="cuda")
    rows[6] = False                                         # the 1e12-scaled row is compared relative to itself
    assert float((a.grad.double() - b.grad)[rows].abs().max()) <= 1e-5 * float(b.grad[rows].abs().max())
    assert float((a.grad.double() - b.grad)[6].abs().max()) <= 1e-5 * float(b.grad[6].abs().max())
    with torch.no_grad():
        assert torch.equal(JF.row_normalize(x), ya.detach())


def test_concat_linear_matches_cat_mm():
    """cat(xs, 1) @ W as a sum of products accumulated in GEMM epilogues (the name / alignment channel of forward_name)."""
    from jmac_b200 import functional as JF
    torch.manual_seed(0)
    N, d = 6000, 256
    for widths in ((256, 256), (256, 256, 256), (128, 256)):
        xs = [torch.randn(N, k, device="cuda", requires_grad=True) for k in widths]
        W = (torch.randn(sum(widths), d, device="cuda") * 0.1).requires_grad_(True)
        G = torch.randn(N, d, device="cuda")
        out = JF.concat_linear(xs, W)
        out.backward(G)
        xr = [x.detach().double().requires_grad_(True) for x in xs]
        Wr = W.detach().double().requires_grad_(True)
        ref = torch.cat(xr, 1) @ Wr
        ref.backward(G.double())
        assert relerr(out.detach(), ref.detach()) < 5e-6
        assert relerr(W.grad, Wr.grad) < 1e-5
        for x, r in zip(xs, xr):
            assert relerr(x.grad, r.grad) < 5e-6
    # small tables: the library path, same values as the plain expression
    xs = [torch.randn(100, 64, device="cuda"), torch.randn(100, 300, device="cuda")]
    W = torch.randn(364, 64, device="cuda")
    assert torch.equal(JF.concat_linear(xs, W), torch.mm(torch.cat(xs, 1), W))
