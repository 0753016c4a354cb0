"""tcgen05 GEMM through the C ABI vs a plain PyTorch fp32 reference of the same op (bf16-rounded inputs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _err_report(got, ref, tag):
    d = (got.float() - ref.float()).abs()
    tol = 2e-2 * ref.float().abs().max().item() + 1e-3
    bad = d > tol
    msg = "%s: max err %.4g (tol %.4g), bad %d / %d" % (tag, d.max().item(), tol, int(bad.sum()), bad.numel())
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        msg += " | bad rows [%d..%d] n=%d, bad cols [%d..%d] n=%d" % (rows.min(), rows.max(), rows.numel(), cols.min(),
                                                                     cols.max(), cols.numel())
    return bool(bad.any()), msg


SHAPES = [(128, 128, 64), (128, 64, 128), (256, 256, 256), (384, 512, 192), (1000, 2304, 768), (130, 72, 200),
          (4096, 128, 2048), (77, 3072, 768), (512, 768, 3072)]


@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_layouts(M, N, K, a_t, b_t):
    from passl_b200 import kernels as K_
    torch.manual_seed(M * 7 + N * 3 + K)
    if (a_t and M % 8) or (b_t and N % 8) or K % 8:
        pytest.skip("leading dims must be multiples of 8")
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    ref = a.float() @ b.float().t()
    aa = a.t().contiguous() if a_t else a
    bb = b.t().contiguous() if b_t else b
    out = K_.gemm(aa, bb, a_t=a_t, b_t=b_t, out_dtype=torch.float32)
    torch.cuda.synchronize()
    bad, msg = _err_report(out, ref, "gemm %s a_t=%s b_t=%s" % ((M, N, K), a_t, b_t))
    assert not bad, msg


def test_gemm_epilogue_bias_gelu_residual_bf16():
    from passl_b200 import kernels as K_
    torch.manual_seed(0)
    M, N, K = 788, 3072, 768
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    ref = torch.nn.functional.gelu(a.float() @ b.float().t() * 0.5 + bias) + res.float()
    out = K_.gemm(a, b, bias=bias, residual=res, act="gelu", alpha=0.5)
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16
    bad, msg = _err_report(out, ref, "epilogue")
    assert not bad, msg


def test_gemm_splitk_atomic_and_colstats():
    from passl_b200 import kernels as K_
    torch.manual_seed(1)
    M, N, K = 256, 192, 8192
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    ref = a.float() @ b.float().t()
    out = K_.gemm(a, b, out_dtype=torch.float32, splits=8)
    torch.cuda.synchronize()
    bad, msg = _err_report(out, ref, "splitk")
    assert not bad, msg
    part = K_.stats_buffer(N, "cuda")                      # per-CTA partials [rows, 2, N], zeroed by the call
    part.fill_(7.0)
    out2 = K_.gemm(a, b, col_stats=part)
    torch.cuda.synchronize()
    o = out2.float()
    cs, cq = part[:, 0].sum(0), part[:, 1].sum(0)
    assert torch.allclose(cs, o.sum(0), rtol=1e-3, atol=1e-1), (cs - o.sum(0)).abs().max()
    assert torch.allclose(cq, (o * o).sum(0), rtol=1e-3, atol=1.0), (cq - (o * o).sum(0)).abs().max()
