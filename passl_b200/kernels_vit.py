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
    _lib.check(lib.passl_b200_bn_bwd_finalize(_ptr(part), nblk, _ptr(sums), _ptr(dgamma), _ptr(dbeta), None, None, None, 0, None, D,
                                              _stream()), "ln_bwd_finalize")
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
    delta = torch.empty((B, H, N), dtype=torch.float32, device=qkv.device)      # workspace of the dO.O pre-pass
    _lib.check(lib.passl_b200_attention_bwd(_ptr(qkv), _ptr(dout.contiguous()), _ptr(out), _ptr(lse), _ptr(dqkv), _ptr(delta), B, N, H, d,
                                            scale, int(causal), _stream()), "attention_bwd")
    return dqkv


# ------------------------------------------------------------------------------------------------------------
# MAE: masking, token assembly, masked-patch MSE
# ------------------------------------------------------------------------------------------------------------
def mae_random_masking(noise, len_keep):
    """noise fp32 [B, L] -> (ids_shuffle int64 [B,L], ids_restore int64 [B,L], mask fp32 [B,L])  (mae.py:184-212)."""
    _need_cuda(noise)
    lib = _lib.load()
    B, L = noise.shape
    assert noise.dtype == torch.float32 and noise.is_contiguous()
    ids_shuffle = torch.empty((B, L), dtype=torch.int64, device=noise.device)
    ids_restore = torch.empty((B, L), dtype=torch.int64, device=noise.device)
    mask = torch.empty((B, L), dtype=torch.float32, device=noise.device)
    _lib.check(lib.passl_b200_mae_random_masking(_ptr(noise), _ptr(ids_shuffle), _ptr(ids_restore), _ptr(mask), B, L, int(len_keep),
                                                 _stream()), "mae_random_masking")
    return ids_shuffle, ids_restore, mask


TOKEN_MODE = {"vit": 0, "mae_enc": 1, "mae_dec": 2}


def token_assemble_fwd(src, pos, tok, B, Ls, Lo, D, mode, ids=None, keep=0):
    lib = _lib.load()
    out = torch.empty((B * Lo, D), dtype=torch.bfloat16, device=src.device)
    _lib.check(lib.passl_b200_token_assemble_fwd(_ptr(src), _ptr(ids), _ptr(pos), _ptr(tok), _ptr(out), B, Ls, Lo, D,
                                                 TOKEN_MODE[mode], int(keep), _stream()), "token_assemble_fwd")
    return out


def token_assemble_bwd(dout, B, Ls, Lo, D, mode, ids=None, keep=0, acc_tok=None, acc_pos=None, ids_tok=None, need_dsrc=True):
    lib = _lib.load()
    dsrc = torch.empty((B * Ls, D), dtype=torch.bfloat16, device=dout.device) if need_dsrc else None
    _lib.check(lib.passl_b200_token_assemble_bwd(_ptr(dout), _ptr(ids), _ptr(dsrc), _ptr(acc_tok), _ptr(acc_pos), _ptr(ids_tok), B, Ls,
                                                 Lo, D, TOKEN_MODE[mode], int(keep), _stream()), "token_assemble_bwd")
    return dsrc


def mae_loss_fwd(pred, imgs, mask, B, Hp, P, pred_tokens, pred_off, norm_pix, mask_sum):
    lib = _lib.load()
    loss = torch.empty(1, dtype=torch.float32, device=pred.device)
    ws = torch.empty(lib.passl_b200_mae_loss_workspace_bytes(B, Hp * Hp), dtype=torch.uint8, device=pred.device)
    _lib.check(lib.passl_b200_mae_loss_fwd(_ptr(pred), _ptr(imgs), _ptr(mask), _ptr(loss), B, Hp, P, pred_tokens, pred_off,
                                           int(norm_pix), float(mask_sum), _ptr(ws), _stream()), "mae_loss_fwd")
    return loss


def mae_loss_bwd(pred, imgs, mask, dloss, B, Hp, P, pred_tokens, pred_off, norm_pix, mask_sum):
    lib = _lib.load()
    dpred = torch.empty_like(pred)
    _lib.check(lib.passl_b200_mae_loss_bwd(_ptr(pred), _ptr(imgs), _ptr(mask), _ptr(dloss), _ptr(dpred), B, Hp, P, pred_tokens,
                                           pred_off, int(norm_pix), float(mask_sum), _stream()), "mae_loss_bwd")
    return dpred


# CLIP: token embedding, EOT pooling, symmetric cross entropy
# ------------------------------------------------------------------------------------------------------------
def embedding_fwd(ids, table, pos):
    """ids int64 [B, L], table fp32 [V, D], pos fp32 [L, D] -> bf16 [B*L, D] = table[ids] + pos  (clip.py:300-303)."""
    _need_cuda(ids, table)
    lib = _lib.load()
    B, L = ids.shape
    V, D = table.shape
    assert ids.dtype == torch.int64 and ids.is_contiguous() and table.dtype == torch.float32 and pos.shape == (L, D)
    out = torch.empty((B * L, D), dtype=torch.bfloat16, device=table.device)
    _lib.check(lib.passl_b200_embedding_fwd(_ptr(ids), _ptr(table), _ptr(pos), _ptr(out), B * L, L, D, V, _stream()), "embedding_fwd")
    return out


def embedding_bwd(ids, dout, dtable=None, dpos=None):
    """Accumulates d table (fp32 [V, D]) and d pos (fp32 [L, D]) from dout bf16 [B*L, D]."""
    lib = _lib.load()
    B, L = ids.shape
    D = dout.shape[1]
    V = dtable.shape[0] if dtable is not None else 1 << 30
    _lib.check(lib.passl_b200_embedding_bwd(_ptr(ids), _ptr(dout), _ptr(dtable), _ptr(dpos), B * L, L, D, V, _stream()),
               "embedding_bwd")


def eot_gather_fwd(ids, x):
    """ids int64 [B, L], x bf16 [B*L, D] -> (x[b, argmax(ids[b])] bf16 [B, D], idx int32 [B])  (clip.py:307-311)."""
    lib = _lib.load()
    B, L = ids.shape
    D = x.shape[1]
    out = torch.empty((B, D), dtype=torch.bfloat16, device=x.device)
    idx = torch.empty(B, dtype=torch.int32, device=x.device)
    _lib.check(lib.passl_b200_eot_gather_fwd(_ptr(ids), _ptr(x), _ptr(out), _ptr(idx), B, L, D, _stream()), "eot_gather_fwd")
    return out, idx


def eot_gather_bwd(idx, dout, L):
    lib = _lib.load()
    B, D = dout.shape
    dx = torch.empty((B * L, D), dtype=torch.bfloat16, device=dout.device)
    _lib.check(lib.passl_b200_eot_gather_bwd(_ptr(idx), _ptr(dout), _ptr(dx), B, L, D, _stream()), "eot_gather_bwd")
    return dx


def clip_ce_fwd(C, logit_scale, n=None, clamp=True):
    """C fp32 [ld, ld] cosine similarities (top-left n x n valid), logit_scale fp32 [1] (log domain, read + clamped on the device)
    -> (out3 fp32 [3] = img_loss, text_loss, loss; workspace to hand to clip_ce_bwd)."""
    _need_cuda(C)
    lib = _lib.load()
    ld = C.shape[0]
    n = ld if n is None else n
    assert C.dtype == torch.float32 and C.is_contiguous() and C.shape == (ld, ld) and logit_scale.dtype == torch.float32
    wsb = lib.passl_b200_clip_ce_workspace_bytes(ld)
    ws = torch.empty(wsb, dtype=torch.uint8, device=C.device)
    out3 = torch.empty(3, dtype=torch.float32, device=C.device)
    _lib.check(lib.passl_b200_clip_ce_fwd(_ptr(C), _ptr(logit_scale), _ptr(out3), n, ld, int(clamp), _ptr(ws), wsb, _stream()),
               "clip_ce_fwd")
    return out3, ws


def clip_ce_bwd(C, ws, n=None, dloss=None, dlogit_scale=None):
    """-> dC bf16 [ld, ld] (zero outside n x n); accumulates d loss / d logit_scale into dlogit_scale (fp32 [1])."""
    lib = _lib.load()
    ld = C.shape[0]
    n = ld if n is None else n
    dC = torch.empty((ld, ld), dtype=torch.bfloat16, device=C.device)
    _lib.check(lib.passl_b200_clip_ce_bwd(_ptr(C), _ptr(dloss), _ptr(dC), _ptr(dlogit_scale), n, ld, _ptr(ws), ws.numel(), _stream()),
               "clip_ce_bwd")
    return dC
