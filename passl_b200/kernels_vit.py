"""Tensor-level wrappers of the ViT-side C-ABI kernels (LayerNorm, fused attention)."""
import torch

from . import _lib
from .kernels import _ptr, _stream, _need_cuda


def layernorm_fwd(x, gamma, beta, eps=1e-6):
    """x bf16 [T, D] -> (y bf16 [T, D], mean fp32 [T], rstd fp32 [T])."""
    _need_cuda(x)
    lib = _lib.load()
    T, D = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    y = torch.empty_like(x)
    mean = torch.empty(T, dtype=torch.float32, device=x.device)
    rstd = torch.empty(T, dtype=torch.float32, device=x.device)
    _lib.check(lib.passl_b200_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(rstd), T, D, float(eps),
                                            _stream()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(x, dy, gamma, mean, rstd, dgamma=None, dbeta=None, dres=None):
    """Returns dx bf16; accumulates dgamma / dbeta into the given fp32 buffers (or returns sums [2, D] = (dbeta, dgamma))."""
    lib = _lib.load()
    T, D = x.shape
    nblk = lib.passl_b200_layernorm_bwd_blocks(T)
    part = torch.empty((nblk, 2, D), dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x)
    _lib.check(lib.passl_b200_layernorm_bwd(_ptr(x), _ptr(dy.contiguous()), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dres), _ptr(dx),
                                            _ptr(part), T, D, _stream()), "layernorm_bwd")
    sums = torch.empty((2, D), dtype=torch.float32, device=x.device)
    _lib.check(lib.passl_b200_bn_bwd_finalize(_ptr(part), nblk, _ptr(sums), _ptr(dgamma), _ptr(dbeta), D, _stream()),
               "ln_bwd_finalize")
    return dx, sums


def attention_fwd(qkv, B, N, H, d, scale=None, causal=False):
    """qkv bf16 [B*N, 3*H*d] (packed [B,N,3,H,d]) -> (out bf16 [B*N, H*d], lse fp32 [B,H,N])."""
    _need_cuda(qkv)
    lib = _lib.load()
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.numel() == B * N * 3 * H * d
    out = torch.empty((B * N, H * d), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)
    scale = float(d) ** -0.5 if scale is None else float(scale)
    _lib.check(lib.passl_b200_attention_fwd(_ptr(qkv), _ptr(out), _ptr(lse), B, N, H, d, scale, int(causal), _stream()),
               "attention_fwd")
    return out, lse


def attention_bwd(qkv, dout, out, lse, B, N, H, d, scale=None, causal=False):
    lib = _lib.load()
    dqkv = torch.empty_like(qkv)
    scale = float(d) ** -0.5 if scale is None else float(scale)
    _lib.check(lib.passl_b200_attention_bwd(_ptr(qkv), _ptr(dout.contiguous()), _ptr(out), _ptr(lse), _ptr(dqkv), B, N, H, d, scale,
                                            int(causal), _stream()), "attention_bwd")
    return dqkv
