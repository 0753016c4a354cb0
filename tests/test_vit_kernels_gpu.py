"""LayerNorm and fused attention (tcgen05) vs a plain PyTorch fp32/fp64 reference of the same op
(passl/models/vision_transformer.py:142-156: q@k^T*scale -> softmax -> @v)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("T,D", [(1000, 768), (197 * 4, 512), (333, 1024), (64, 1536)])
def test_layernorm_fwd_bwd(T, D):
    from passl_b200 import kernels_vit as V
    torch.manual_seed(0)
    x = (torch.randn(T, D, device="cuda") * 2 + 0.3).bfloat16()
    g = torch.rand(D, device="cuda") + 0.5
    b = torch.randn(D, device="cuda")
    y, mean, rstd = V.layernorm_fwd(x, g, b, eps=1e-6)
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), gr, br, eps=1e-6)
    assert rel(y, yr) < 4e-3
    dy = torch.randn(T, D, device="cuda").bfloat16()
    yr.backward(dy.double())
    dg = torch.zeros(D, device="cuda")
    db = torch.zeros(D, device="cuda")
    dx, sums = V.layernorm_bwd(x, dy, g, mean, rstd, dgamma=dg, dbeta=db)
    torch.cuda.synchronize()
    assert rel(dx, xr.grad) < 6e-3
    assert rel(dg, gr.grad) < 1e-3 and rel(db, br.grad) < 1e-3
    assert torch.equal(sums[0], db) and torch.equal(sums[1], dg)


def _ref_attention(qkv, B, N, H, d, causal):
    x = qkv.double().reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)      # [3, B, H, N, d]
    q, k, v = x[0], x[1], x[2]
    s = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    if causal:
        s = s + torch.full((N, N), float("-inf"), dtype=s.dtype, device=s.device).triu(1)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * N, H * d)
    return o, torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("B,N,H,d,causal", [(3, 197, 12, 64, False), (2, 50, 12, 64, False), (2, 197, 16, 32, False),
                                            (4, 77, 8, 64, True), (2, 256, 4, 64, False), (1, 128, 2, 32, True),
                                            (40, 197, 12, 64, False)])
def test_attention_fwd_bwd(B, N, H, d, causal):
    from passl_b200 import kernels_vit as V
    torch.manual_seed(B * 1000 + N)
    qkv = torch.randn(B * N, 3 * H * d, device="cuda").bfloat16()
    out, lse = V.attention_fwd(qkv, B, N, H, d, causal=causal)
    qr = qkv.double().requires_grad_(True)
    oref, lref = _ref_attention(qr, B, N, H, d, causal)
    torch.cuda.synchronize()
    assert rel(out, oref) < 1e-2, rel(out, oref)
    assert torch.allclose(lse.double(), lref, rtol=1e-4, atol=1e-3), (lse.double() - lref).abs().max()
    dout = torch.randn_like(out)
    oref.backward(dout.double())
    dqkv = V.attention_bwd(qkv, dout, out, lse, B, N, H, d, causal=causal)
    torch.cuda.synchronize()
    g = qr.grad.reshape(B, N, 3, H, d)
    got = dqkv.double().reshape(B, N, 3, H, d)
    for s_, name in enumerate("qkv"):
        r = rel(got[:, :, s_], g[:, :, s_])
        assert r < 2e-2, (name, r)
