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


def _gelu_grad(u):
    u = u.double()
    return (0.5 * (1 + torch.erf(u / 2 ** 0.5)) + u * torch.exp(-u * u / 2) / (2 * torch.pi) ** 0.5).float()


@pytest.mark.parametrize("M,N,K", [(777, 200, 136), (128, 32, 64), (1000, 2304, 768), (33, 3072, 192)])
def test_gemm_linear_epilogue_tma_store(M, N, K):
    """bf16 row-major outputs of the linear layers (gemm.cuh EPI 1: swizzled 32 x 32 boxes leave through TMA stores, ragged edges
    are clipped by the tensor map): bias, GELU + saved pre-activation, gates from a saved pre-activation, residual."""
    from passl_b200 import kernels as K_
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    y = a.float() @ w.float().t()
    out = K_.gemm(a, w, bias=bias)
    bad, msg = _err_report(out, y + bias, "bias")
    assert out.dtype == torch.bfloat16 and not bad, msg
    u = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    out = K_.gemm(a, w, bias=bias, act="gelu", preact_out=u)
    bad, msg = _err_report(u, y + bias, "pre-activation")
    assert not bad, msg
    bad, msg = _err_report(out, torch.nn.functional.gelu(y + bias), "gelu")
    assert not bad, msg
    out = K_.gemm(a, w, bias=bias, act="quick_gelu")
    bad, msg = _err_report(out, (y + bias) * torch.sigmoid(1.702 * (y + bias)), "quick_gelu")
    assert not bad, msg
    res = torch.randn(M, N, device="cuda").bfloat16()
    out = K_.gemm(a, w, bias=bias, residual=res, act="relu")          # residual through the epilogue (an activation precedes it)
    bad, msg = _err_report(out, torch.relu(y + bias) + res.float(), "relu + residual")
    assert not bad, msg
    out = K_.gemm(a, w, bias=bias, residual=res)                      # residual through the MMA when the shape allows
    bad, msg = _err_report(out, y + bias + res.float(), "residual")
    assert not bad, msg
    aux = torch.randn(M, N, device="cuda").bfloat16()
    for mode, gate in [("relu_mask", (aux.float() > 0).float()), ("gelu_grad", _gelu_grad(aux)),
                       ("quick_gelu_grad", (lambda s: s * (1 + 1.702 * aux.float() * (1 - s)))(torch.sigmoid(1.702 * aux.float())))]:
        out = K_.gemm(a, w, aux=aux, aux_mode_name=mode)
        bad, msg = _err_report(out, y * gate, mode)
        assert not bad, msg
    # dgrad form (B given as [K, N]) with the GELU' gate, as fc2's backward issues it
    wt = w.t().contiguous()
    out = K_.gemm(a, wt, b_t=True, aux=aux, aux_mode_name="gelu_grad")
    bad, msg = _err_report(out, y * _gelu_grad(aux), "gelu_grad, b_t")
    assert not bad, msg


def test_gelu_gate_arithmetic_is_exact_to_bf16():
    """The A&S erfc forms in the epilogue against erf in double: the gate / activation error must vanish under bf16 rounding."""
    from passl_b200 import kernels as K_
    M, N, K = 256, 256, 64
    a = torch.zeros(M, K, device="cuda", dtype=torch.bfloat16)
    a[:, 0] = 1.0
    w = torch.zeros(N, K, device="cuda", dtype=torch.bfloat16)
    w[:, 0] = 1.0                                                  # a @ w.T = 1 everywhere: out = gate(aux)
    aux = torch.linspace(-8, 8, M * N, device="cuda").view(M, N).bfloat16()
    out = K_.gemm(a, w, aux=aux, aux_mode_name="gelu_grad").float()
    ref = _gelu_grad(aux)
    assert (out - ref).abs().max().item() < 6e-3                   # bf16 rounding of values in [-0.13, 1.13] is <= 3.9e-3
    assert (out - ref.bfloat16().float()).abs().mean().item() < 2e-4
    bias = torch.linspace(-6, 6, N, device="cuda")                 # activation of a row of pre-activations
    a0 = torch.zeros(M, K, device="cuda", dtype=torch.bfloat16)
    out = K_.gemm(a0, w, bias=bias, act="gelu").float()
    ref = torch.nn.functional.gelu(bias.double()).float().expand(M, N)
    assert bool(((out - ref).abs() <= 2 ** -8 * ref.abs() + 2e-5).all()), (out - ref).abs().max().item()


@pytest.mark.parametrize("b_t", [False, True])
def test_gemm_cta_pair(b_t):
    """Shapes big enough for the CTA-pair path (cluster of 2, tcgen05 cta_group::2, gemm.cuh CG = 2): odd number of 128-row
    blocks (the last pair has one CTA outside the tensor), fp32 and bf16 outputs, bias, column statistics, residual."""
    from passl_b200 import kernels as K_
    torch.manual_seed(5)
    M, N, K = 128 * 150 + 40, 512, 512
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    ww = w.t().contiguous() if b_t else w
    y = a.float() @ w.float().t()
    out = K_.gemm(a, ww, b_t=b_t, out_dtype=torch.float32)
    bad, msg = _err_report(out, y, "pair fp32")
    assert not bad, msg
    bias = torch.randn(N, device="cuda")
    out = K_.gemm(a, ww, b_t=b_t, bias=bias)
    bad, msg = _err_report(out, y + bias, "pair bf16 + bias")
    assert not bad, msg
    part = K_.stats_buffer(N, "cuda")
    out2 = K_.gemm(a, ww, b_t=b_t, col_stats=part)
    o = out2.float()
    cs, cq = part[:, 0].sum(0), part[:, 1].sum(0)
    assert torch.allclose(cs, o.sum(0), rtol=1e-3, atol=0.5), (cs - o.sum(0)).abs().max()
    assert torch.allclose(cq, (o * o).sum(0), rtol=1e-3, atol=2.0), (cq - (o * o).sum(0)).abs().max()
    # long K loop + residual: the pair takes the residual through the epilogue's operand tile
    K2 = 1536
    a2 = torch.randn(M, K2, device="cuda").bfloat16()
    w2 = (torch.randn(N, K2, device="cuda") / K2 ** 0.5).bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16()
    out = K_.gemm(a2, w2.t().contiguous() if b_t else w2, b_t=b_t, bias=bias, residual=res)
    bad, msg = _err_report(out, a2.float() @ w2.float().t() + bias + res.float(), "pair + residual")
    assert not bad, msg


@pytest.mark.parametrize("K", [128, 256, 768])
def test_gemm_sixteen_epilogue_warps(K):
    """Large-M launches on 256-wide tiles that take the 16-epilogue-warp instantiations (gemm.cuh EW = 16): GELU + saved
    pre-activation (K < 768 single CTA, K >= 768 CTA pair), the GELU' / QuickGELU' gates in both weight layouts, and the K-small
    1x1-convolution form with BatchNorm column statistics.  M is not a multiple of 128 and the number of row blocks is odd."""
    from passl_b200 import kernels as K_
    torch.manual_seed(K)
    M, N = 128 * 150 + 40, 512
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    y = a.float() @ w.float().t()
    u = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out = K_.gemm(a, w, bias=bias, act="gelu", preact_out=u)
    bad, msg = _err_report(u, y + bias, "pre-activation")
    assert not bad, msg
    bad, msg = _err_report(out, torch.nn.functional.gelu(y + bias), "gelu")
    assert not bad, msg
    out = K_.gemm(a, w, bias=bias, act="quick_gelu")
    bad, msg = _err_report(out, (y + bias) * torch.sigmoid(1.702 * (y + bias)), "quick_gelu")
    assert not bad, msg
    aux = torch.randn(M, N, device="cuda").bfloat16()
    sg = torch.sigmoid(1.702 * aux.float())
    for b_t in (False, True):
        ww = w.t().contiguous() if b_t else w
        out = K_.gemm(a, ww, b_t=b_t, aux=aux, aux_mode_name="gelu_grad")
        bad, msg = _err_report(out, y * _gelu_grad(aux), "gelu_grad b_t=%s" % b_t)
        assert not bad, msg
        out = K_.gemm(a, ww, b_t=b_t, aux=aux, aux_mode_name="quick_gelu_grad")
        bad, msg = _err_report(out, y * (sg * (1 + 1.702 * aux.float() * (1 - sg))), "quick_gelu_grad b_t=%s" % b_t)
        assert not bad, msg
    part = K_.stats_buffer(N, "cuda")
    out2 = K_.gemm(a, w, col_stats=part)                     # 1x1-convolution form: statistics of the stored bf16 values
    bad, msg = _err_report(out2, y, "stats output")
    assert not bad, msg
    o = out2.float()
    cs, cq = part[:, 0].sum(0), part[:, 1].sum(0)
    assert torch.allclose(cs, o.sum(0), rtol=1e-3, atol=0.5), (cs - o.sum(0)).abs().max()
    assert torch.allclose(cq, (o * o).sum(0), rtol=1e-3, atol=2.0), (cq - (o * o).sum(0)).abs().max()
