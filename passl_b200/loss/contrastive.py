"""Contrastive losses as single autograd nodes over the fused kernels (the new home SURVEY.md §1 fact 1 allows for
"passl/loss/contrastive": the reference computes these inside model methods / heads).

  moco_infonce   : moco.py:178-182 + contrastive_head.py:37-60     [l_pos | l_neg]/T, labels 0, mean CE, top-1/5
  gathered_infonce: mocov3.py:187-198 / clip.py:331-335             q . k_all^T * scale, labels arange(N)+N*rank

Forward never materialises the logits (tcgen05 kernel streams the key matrix once); backward recomputes them tile by
tile.  Only the queries receive a gradient (keys / queue are no-grad in the reference: moco.py:162-180,
mocov3.py:173-198).
"""
import torch

from .. import kernels as K


class _FusedInfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, keys, pos, label, scale, loss_scale, precision):
        # q fp32 [N, D] (normalised), keys [K, D] bf16 or fp32, pos fp32 [N, D] or None, label int64 [N] or None
        q = q.contiguous()
        if precision == "bf16":
            qb = K.cast_bf16(q)
            kb = keys if keys.dtype == torch.bfloat16 else K.cast_bf16(keys)
            out, lse, tgt, _ = K.infonce_tc_fwd(qb, kb, pos=pos, label=label, scale=scale, loss_scale=loss_scale)
            ctx.q_bwd, ctx.keys_bwd = qb, kb
        else:
            out, lse, tgt, _ = K.simce_fwd(q, keys, pos=pos, label=label, scale=scale, loss_scale=loss_scale)
            ctx.q_bwd, ctx.keys_bwd = q, keys
        ctx.pos, ctx.label, ctx.lse, ctx.tgt = pos, label, lse, tgt
        ctx.scale, ctx.loss_scale, ctx.precision = scale, loss_scale, precision
        ctx.mark_non_differentiable(out[1], out[2])
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, dloss, _a1, _a5):
        qa = ctx.q_bwd
        dl = dloss.contiguous().float()
        if qa.dtype == torch.bfloat16 and qa.shape[1] <= 256:
            # tcgen05 backward: S recomputed tile by tile, P (bf16) x key tile accumulated into dQ in TMEM
            dq = K.infonce_tc_bwd(qa, ctx.keys_bwd, ctx.lse, ctx.tgt, pos=ctx.pos, label=ctx.label, scale=ctx.scale,
                                  loss_scale=ctx.loss_scale, dloss=dl)
            return dq, None, None, None, None, None, None
        if qa.dtype != torch.float32:        # the SIMT backward kernel takes fp32 queries (bf16-rounded values)
            qf = torch.empty(qa.shape, dtype=torch.float32, device=qa.device)
            K.cast_f32(qa, qf)
            qa = qf
        dq = K.simce_bwd(qa, ctx.keys_bwd, ctx.lse, ctx.tgt, pos=ctx.pos, label=ctx.label, scale=ctx.scale,
                         loss_scale=ctx.loss_scale, dloss=dl)
        return dq, None, None, None, None, None, None


def moco_infonce(q, k, queue, temperature, precision="bf16"):
    """q, k: fp32 [N, D] L2-normalised; queue: key-major [K, D].  Returns (loss, acc1, acc5) device scalars."""
    return _FusedInfoNCE.apply(q, queue, k.detach().contiguous(), None, 1.0 / temperature, 1.0, precision)


def gathered_infonce(q, k_all, labels, scale, loss_scale=1.0, precision="bf16"):
    """q fp32 [N, D]; k_all [M, D] gathered keys (no grad); labels int64 [N].  Returns (loss, acc1, acc5)."""
    return _FusedInfoNCE.apply(q, k_all.detach().contiguous(), None, labels, scale, loss_scale, precision)


class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode, eps):
        y, _, inv = K.l2norm_fwd(x.contiguous(), mode=mode, eps=eps)
        ctx.y, ctx.inv, ctx.mode, ctx.eps = y, inv, mode, eps
        return y

    @staticmethod
    def backward(ctx, dy):
        dx, _ = K.l2norm_bwd(dy.contiguous(), ctx.y, ctx.inv, mode=ctx.mode, eps=ctx.eps)
        return dx, None, None


def normalize(x, eps=1e-12):
    """paddle.nn.functional.normalize(x, axis=1): x / max(||x||, eps)."""
    return _L2Normalize.apply(x, "normalize", eps)


def l2_normalize(x, eps=1e-12):
    """passl/nn/norm.py:18-40: x / sqrt(sum x^2 + eps)."""
    return _L2Normalize.apply(x, "l2_normalize", eps)


class _LogitsCE(torch.autograd.Function):
    """nn.CrossEntropyLoss()(logits, labels) on materialised logits (reference head signatures, clip_head.py:29-32)."""

    @staticmethod
    def forward(ctx, logits, labels):
        from .. import _lib
        lib = _lib.load()
        x = logits.detach().float().contiguous()
        n, m = x.shape
        lb = labels.to(torch.int64).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        lse = torch.empty(n, dtype=torch.float32, device=x.device)
        ws = torch.empty(n, dtype=torch.float32, device=x.device)
        _lib.check(lib.passl_b200_rows_ce_fwd(K._ptr(x), K._ptr(lb), K._ptr(loss), K._ptr(lse), n, m, K._ptr(ws), n * 4, K._stream()),
                   "rows_ce_fwd")
        ctx.saved = (x, lb, lse)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        from .. import _lib
        lib = _lib.load()
        x, lb, lse = ctx.saved
        dx = torch.empty_like(x)
        _lib.check(lib.passl_b200_rows_ce_bwd(K._ptr(x), K._ptr(lb), K._ptr(lse), K._ptr(dloss.contiguous().float().reshape(1)),
                                              K._ptr(dx), x.shape[0], x.shape[1], K._stream()), "rows_ce_bwd")
        return dx, None


def logits_cross_entropy(logits, labels):
    K._need_cuda(logits)
    return _LogitsCE.apply(logits, labels)
