"""Thin tensor-level wrappers over the C ABI: torch is only the carrier of device pointers and the current stream."""
import torch

from . import _lib

ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2, "quick_gelu": 3}


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.PasslB200Error("passl_b200 kernels need CUDA tensors (there is no CPU fallback)")


_ws_cache = {}


def workspace(nbytes, device, tag="default"):
    key = (tag, str(device))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ------------------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------------------
def gemm(a, b, *, a_t=False, b_t=False, out=None, out_dtype=torch.bfloat16, bias=None, residual=None, act=None,
         alpha=1.0, splits=1, accumulate=False, col_stats=None, aux=None, aux_mode_name="relu_mask", preact_out=None):
    """out[M,N] = act(alpha * A @ B^T + bias) + residual.

    a: [M,K] (or [K,M] when a_t), b: [N,K] (or [K,N] when b_t), bf16, last dim contiguous.
    """
    _need_cuda(a, b)
    lib = _lib.load()
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    if a_t:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_t:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape)
    if out is None:
        out = (torch.zeros if (accumulate or splits > 1) else torch.empty)((M, N), dtype=out_dtype, device=a.device)
    out_fp32 = out.dtype == torch.float32
    atomic = 1 if (accumulate or splits > 1) else 0
    cs, cq = col_stats, None          # col_stats: fp32 [gemm_stats_rows(), 2, N] partials buffer (see stats_buffer)
    aux_mode = 0
    if aux is not None:
        aux_mode = {"relu_mask": 1, "gelu_grad": 2, "quick_gelu_grad": 3}[aux_mode_name]
        assert aux.dtype == torch.bfloat16 and aux.stride(0) == out.stride(0)
    code = lib.passl_b200_gemm_bf16_ex(_ptr(a), _ptr(b), _ptr(out), M, N, K, int(a_t), int(b_t), a.stride(0), b.stride(0),
                                       out.stride(0), int(out_fp32), atomic, _ptr(bias), _ptr(residual), ACT[act],
                                       float(alpha), int(splits), _ptr(cs), _ptr(cq), _ptr(aux), aux_mode, _ptr(preact_out),
                                       _stream())
    _lib.check(code, "gemm_bf16")
    return out


def stats_buffer(C, device):
    """Partials buffer for BatchNorm statistics fused into the producing GEMM / conv epilogue: [rows, 2, C] fp32."""
    rows = _lib.load().passl_b200_gemm_stats_rows()
    return torch.empty((rows, 2, C), dtype=torch.float32, device=device)


def wgrad_splits(M, N, K):
    """Split-K factor for weight-gradient GEMMs (small M x N output, long K)."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    s = max(1, (2 * 148 + tiles - 1) // tiles)
    return max(1, min(s, (K + 63) // 64 // 4))


# ------------------------------------------------------------------------------------------------------------
# Convolution (NHWC bf16, weights [Cout,R,S,Cin])
# ------------------------------------------------------------------------------------------------------------
def conv2d_fwd(x, w, stride=1, pad=0, bias=None, residual=None, act=None, col_stats=None, out=None):
    _need_cuda(x, w)
    lib = _lib.load()
    N, H, W, Cin = x.shape
    Cout, R, S, Cin2 = w.shape
    assert Cin == Cin2 and x.is_contiguous() and w.is_contiguous()
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W + 2 * pad - S) // stride + 1
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
    cs, cq = col_stats, None          # fp32 [gemm_stats_rows(), 2, Cout] partials buffer (see stats_buffer)
    code = lib.passl_b200_conv2d_fwd_bf16(_ptr(x), _ptr(w), _ptr(out), N, H, W, Cin, Cout, R, S, stride, pad, _ptr(bias),
                                          _ptr(residual), ACT[act], _ptr(cs), _ptr(cq), _stream())
    _lib.check(code, "conv2d_fwd_bf16")
    return out


def conv2d_dgrad(dy, w, x_shape, stride=1, pad=0, out=None, accumulate=False):
    _need_cuda(dy, w)
    lib = _lib.load()
    N, H, W, Cin = x_shape
    Cout, R, S, _ = w.shape
    assert dy.is_contiguous() and w.is_contiguous()
    if out is None:
        assert not accumulate
        out = torch.empty((N, H, W, Cin), dtype=torch.bfloat16, device=dy.device)
    ws = workspace(lib.passl_b200_conv2d_dgrad_workspace_bytes(Cin, Cout, R, S), dy.device, "dgrad")
    code = lib.passl_b200_conv2d_dgrad_bf16(_ptr(dy), _ptr(w), _ptr(out), _ptr(ws), N, H, W, Cin, Cout, R, S, stride, pad,
                                            int(accumulate), _stream())
    _lib.check(code, "conv2d_dgrad_bf16")
    return out


def conv2d_wgrad(x, dy, w_shape, stride=1, pad=0, out=None, accumulate=False):
    _need_cuda(x, dy)
    lib = _lib.load()
    N, H, W, Cin = x.shape
    Cout, R, S, _ = w_shape
    assert x.is_contiguous() and dy.is_contiguous()
    if out is None:
        out = torch.empty((Cout, R, S, Cin), dtype=torch.float32, device=x.device)
        accumulate = False
    code = lib.passl_b200_conv2d_wgrad_bf16(_ptr(x), _ptr(dy), _ptr(out), N, H, W, Cin, Cout, R, S, stride, pad,
                                            int(not accumulate), _stream())
    _lib.check(code, "conv2d_wgrad_bf16")
    return out


# ------------------------------------------------------------------------------------------------------------
# ResNet stem (7x7/2, 3 -> 64) as a 4x1 convolution over the W-unfolded space-to-depth repack (csrc/stem.cu)
# ------------------------------------------------------------------------------------------------------------
def stem_pack_input(img):
    """NCHW fp32 [N,3,H,W] -> bf16 [N,H/2,W/2,64]."""
    _need_cuda(img)
    lib = _lib.load()
    N, C, H, W = img.shape
    assert C == 3 and img.dtype == torch.float32 and img.is_contiguous()
    xp = torch.empty((N, H // 2, W // 2, 64), dtype=torch.bfloat16, device=img.device)
    _lib.check(lib.passl_b200_stem_pack_input(_ptr(img), _ptr(xp), N, H, W, _stream()), "stem_pack_input")
    return xp


def stem_pack_weight(w):
    """fp32 [64, kpad] ((r,s,c) order) -> bf16 [64,4,1,64]."""
    lib = _lib.load()
    wp = torch.empty((64, 4, 1, 64), dtype=torch.bfloat16, device=w.device)
    _lib.check(lib.passl_b200_stem_pack_weight(_ptr(w), _ptr(wp), w.shape[1], _stream()), "stem_pack_weight")
    return wp


def stem_conv_fwd(xp, wp, col_stats=None):
    lib = _lib.load()
    N, H, W, _ = xp.shape
    out = torch.empty((N, H, W, 64), dtype=torch.bfloat16, device=xp.device)
    _lib.check(lib.passl_b200_conv2d_fwd_rect_bf16(_ptr(xp), _ptr(wp), _ptr(out), N, H, W, 64, 64, 4, 1, 2, 0, H, W, None, 0,
                                                   _ptr(col_stats), _stream()), "conv2d_fwd_rect_bf16")
    return out


def stem_conv_wgrad(xp, dy, dw_acc):
    """dw_acc fp32 [64, kpad] += d/dw of the stem for dy bf16 [N,H/2,W/2,64]."""
    lib = _lib.load()
    N, H, W, _ = xp.shape
    dwp = torch.empty((64, 4, 1, 64), dtype=torch.float32, device=xp.device)
    _lib.check(lib.passl_b200_conv2d_wgrad_rect_bf16(_ptr(xp), _ptr(dy), _ptr(dwp), N, H, W, 64, 64, 4, 1, 2, 0, H, W, 1, _stream()),
               "conv2d_wgrad_rect_bf16")
    _lib.check(lib.passl_b200_stem_unpack_wgrad(_ptr(dwp), _ptr(dw_acc), dw_acc.shape[1], _stream()), "stem_unpack_wgrad")


# ------------------------------------------------------------------------------------------------------------
# Fused similarity / softmax / CE (fp32 SIMT variant)
# ------------------------------------------------------------------------------------------------------------
def simce_fwd(a, b, *, pos=None, label=None, excl=None, scale=1.0, loss_scale=1.0, want_rows=False):
    _need_cuda(a, b)
    lib = _lib.load()
    N, D = a.shape
    K = b.shape[0]
    assert a.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous()
    lse = torch.empty(N, dtype=torch.float32, device=a.device)
    tgt = torch.empty(N, dtype=torch.float32, device=a.device)
    rows = torch.empty(N, dtype=torch.float32, device=a.device) if want_rows else None
    out = torch.empty(3, dtype=torch.float32, device=a.device)
    nb = lib.passl_b200_simce_workspace_bytes(N, K)
    ws = workspace(nb, a.device, "simce")
    code = lib.passl_b200_simce_fwd_f32(_ptr(a), _ptr(b), int(b.dtype == torch.bfloat16), _ptr(pos), _ptr(label),
                                        _ptr(excl), float(scale), float(loss_scale), N, K, D, _ptr(lse), _ptr(tgt),
                                        _ptr(rows), _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "simce_fwd_f32")
    return out, lse, tgt, rows


def simce_bwd(a, b, lse, tgt, *, pos=None, label=None, excl=None, scale=1.0, loss_scale=1.0, dloss=None):
    lib = _lib.load()
    N, D = a.shape
    K = b.shape[0]
    da = torch.empty_like(a)
    ws = workspace(N * 4 + 256, a.device, "simce_bwd")
    code = lib.passl_b200_simce_bwd_f32(_ptr(a), _ptr(b), int(b.dtype == torch.bfloat16), _ptr(pos), _ptr(label),
                                        _ptr(excl), float(scale), float(loss_scale), N, K, D, _ptr(lse), _ptr(tgt),
                                        _ptr(dloss), _ptr(da), _ptr(ws), ws.numel(), _stream())
    _lib.check(code, "simce_bwd_f32")
    return da


_nce_state = {}


def _infonce_state(nbytes, device, N):
    """Persistent state of the fused InfoNCE kernel (epoch, ticket, target flags, slice accumulators): zero-filled ONCE per
    (device, stream, N); every launch leaves it ready for the next one (include/passl_b200.h)."""
    key = (str(device), int(torch.cuda.current_stream(device).cuda_stream), int(N))
    buf = _nce_state.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
        _nce_state[key] = buf
    return buf


def infonce_tc_fwd(q_bf16, keys_bf16, *, pos=None, label=None, excl=None, scale=1.0, loss_scale=1.0, want_rows=False):
    """tcgen05 fused InfoNCE forward (one launch): q [N,D] bf16, keys [K,D] bf16 (streamed once), pos [N,D] fp32 optional."""
    _need_cuda(q_bf16, keys_bf16)
    lib = _lib.load()
    N, D = q_bf16.shape
    K = keys_bf16.shape[0]
    assert q_bf16.dtype == torch.bfloat16 and keys_bf16.dtype == torch.bfloat16
    assert q_bf16.is_contiguous() and keys_bf16.is_contiguous()
    dev = q_bf16.device
    lse = torch.empty(N, dtype=torch.float32, device=dev)
    tgt = torch.empty(N, dtype=torch.float32, device=dev)
    rows = torch.empty(N, dtype=torch.float32, device=dev) if want_rows else None
    out = torch.empty(3, dtype=torch.float32, device=dev)
    ws = _infonce_state(lib.passl_b200_infonce_tc_workspace_bytes(N, K, D), dev, N)
    code = lib.passl_b200_infonce_tc_fwd(_ptr(q_bf16), _ptr(keys_bf16), _ptr(pos), _ptr(label), _ptr(excl), float(scale),
                                         float(loss_scale), N, K, D, _ptr(lse), _ptr(tgt), _ptr(rows), _ptr(out), _ptr(ws),
                                         ws.numel(), _stream())
    _lib.check(code, "infonce_tc_fwd")
    return out, lse, tgt, rows


def infonce_tc_bwd(q_bf16, keys_bf16, lse, tgt, *, pos=None, label=None, excl=None, scale=1.0, loss_scale=1.0, dloss=None):
    """tcgen05 fused InfoNCE backward w.r.t. the queries (one launch + the zero fill of dq): returns dq fp32 [N, D]."""
    _need_cuda(q_bf16, keys_bf16)
    lib = _lib.load()
    N, D = q_bf16.shape
    K = keys_bf16.shape[0]
    assert q_bf16.dtype == torch.bfloat16 and keys_bf16.dtype == torch.bfloat16
    assert q_bf16.is_contiguous() and keys_bf16.is_contiguous()
    dq = torch.empty((N, D), dtype=torch.float32, device=q_bf16.device)
    code = lib.passl_b200_infonce_tc_bwd(_ptr(q_bf16), _ptr(keys_bf16), _ptr(pos), _ptr(label), _ptr(excl), float(scale),
                                         float(loss_scale), N, K, D, _ptr(lse), _ptr(tgt), _ptr(dloss), _ptr(dq), _stream())
    _lib.check(code, "infonce_tc_bwd")
    return dq


# ------------------------------------------------------------------------------------------------------------
# Embedding utilities
# ------------------------------------------------------------------------------------------------------------
L2_MODE = {"normalize": 0, "l2_normalize": 1, "clip": 2}


def l2norm_fwd(x, mode="normalize", eps=1e-12, want_bf16=False):
    _need_cuda(x)
    lib = _lib.load()
    N, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty_like(x)
    yb = torch.empty((N, D), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    inv = torch.empty(N, dtype=torch.float32, device=x.device)
    _lib.check(lib.passl_b200_l2norm_fwd(_ptr(x), _ptr(y), _ptr(yb), _ptr(inv), N, D, L2_MODE[mode], float(eps), _stream()),
               "l2norm_fwd")
    return y, yb, inv


def l2norm_bwd(dy, y, inv, mode="normalize", eps=1e-12, want_bf16=False):
    lib = _lib.load()
    N, D = y.shape
    dx = torch.empty_like(y)
    dxb = torch.empty((N, D), dtype=torch.bfloat16, device=y.device) if want_bf16 else None
    _lib.check(lib.passl_b200_l2norm_bwd(_ptr(dy.contiguous()), _ptr(y), _ptr(inv), _ptr(dx), _ptr(dxb), N, D,
                                         L2_MODE[mode], float(eps), _stream()), "l2norm_bwd")
    return dx, dxb


def queue_enqueue(keys, queue_ptr, queue_f32=None, queue_bf16=None):
    """queue[ptr:ptr+Bg] = keys; ptr = (ptr+Bg) % K — all on device (moco.py:92-105)."""
    _need_cuda(keys, queue_ptr)
    lib = _lib.load()
    Bg, D = keys.shape
    q = queue_f32 if queue_f32 is not None else queue_bf16
    K = q.shape[0]
    if K % Bg != 0:
        raise AssertionError("K %% batch_size != 0 (moco.py:99): K=%d batch=%d" % (K, Bg))
    assert queue_ptr.dtype == torch.int64 and keys.dtype == torch.float32 and keys.is_contiguous()
    _lib.check(lib.passl_b200_queue_enqueue(_ptr(keys), _ptr(queue_f32), _ptr(queue_bf16), _ptr(queue_ptr), Bg, D, K,
                                            _stream()), "queue_enqueue")


def ema_update(k, q, m, k_bf16=None):
    _need_cuda(k, q)
    lib = _lib.load()
    assert k.dtype == torch.float32 and q.dtype == torch.float32 and k.numel() == q.numel()
    _lib.check(lib.passl_b200_ema_update(_ptr(k), _ptr(q), _ptr(k_bf16), float(m), k.numel(), _stream()), "ema_update")


def cast_bf16(x, out=None):
    _need_cuda(x)
    lib = _lib.load()
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.passl_b200_cast_f32_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), "cast_f32_to_bf16")
    return out


# ------------------------------------------------------------------------------------------------------------
# BatchNorm (channels-last bf16 [P, C]) and pooling
# ------------------------------------------------------------------------------------------------------------
def bn_stats(y2d):
    """Per-CTA partial sums of y and y^2: fp32 [nblk, 2, C] (summed by bn_finalize; no atomics)."""
    lib = _lib.load()
    P, C = y2d.shape
    nblk = lib.passl_b200_bn_reduce_blocks(P, C)
    part = torch.empty((nblk, 2, C), dtype=torch.float32, device=y2d.device)
    _lib.check(lib.passl_b200_bn_stats(_ptr(y2d), _ptr(part), P, C, _stream()), "bn_stats")
    return part


def bn_finalize(part, gamma, beta, running_mean, running_var, count, eps=1e-5, momentum=0.9):
    """part: fp32 [nblk, 2, C] partials (or [2, C] totals) -> fp32 [4, C] = (mean, invstd, scale, shift); updates the
    running stats in place (may be None)."""
    lib = _lib.load()
    C = part.shape[-1]
    nblk = part.shape[0] if part.dim() == 3 else 1
    out = torch.empty((4, C), dtype=torch.float32, device=part.device)
    _lib.check(lib.passl_b200_bn_finalize(_ptr(part), nblk, _ptr(gamma), _ptr(beta), _ptr(out[0]), _ptr(out[1]),
                                          _ptr(out[2]), _ptr(out[3]), _ptr(running_mean), _ptr(running_var), int(count),
                                          float(eps), float(momentum), C, _stream()), "bn_finalize")
    return out


def colsum_accumulate(x2d, acc):
    """acc[c] += sum_p x2d[p, c]  (bias gradients): reduce kernel partials + the finalize kernel's accumulate path."""
    lib = _lib.load()
    P, C = x2d.shape
    part = bn_stats(x2d)
    sums = torch.empty((2, C), dtype=torch.float32, device=x2d.device)
    _lib.check(lib.passl_b200_bn_bwd_finalize(_ptr(part), part.shape[0], _ptr(sums), None, _ptr(acc), None, None, None, 0, None, C,
                                              _stream()), "bn_bwd_finalize")


def bn_apply(y, msss, relu, residual=None, out=None, out_f32=None):
    lib = _lib.load()
    C = y.shape[-1]
    P = y.numel() // C
    if out is None and out_f32 is None:
        out = torch.empty_like(y)
    _lib.check(lib.passl_b200_bn_apply(_ptr(y), _ptr(residual), _ptr(msss[2]), _ptr(msss[3]), _ptr(out), _ptr(out_f32), P, C,
                                       int(relu), _stream()), "bn_apply")
    return out if out is not None else out_f32


def bn_apply_mask(y, msss, residual=None):
    """z = relu(y*scale + shift + residual) plus the ReLU mask as 1 bit per element (uint8 [P*C/8]) for the backward."""
    lib = _lib.load()
    C = y.shape[-1]
    P = y.numel() // C
    out = torch.empty_like(y)
    mask = torch.empty(P * C // 8, dtype=torch.uint8, device=y.device)
    _lib.check(lib.passl_b200_bn_apply_mask(_ptr(y), _ptr(residual), _ptr(msss[2]), _ptr(msss[3]), _ptr(out), _ptr(mask), P, C, 1,
                                            _stream()), "bn_apply_mask")
    return out, mask


def bn_bwd(y, dz, z, msss, gamma, relu, want_dres=False, dgamma=None, dbeta=None, mask_bits=None):
    """Returns (dy, dres, sums) with sums fp32 [2, C] = (dbeta, dgamma) of this call; when the fp32 gradient buffers
    dgamma / dbeta are given the totals are accumulated into them by the same tiny kernel."""
    lib = _lib.load()
    C = y.shape[-1]
    P = y.numel() // C
    nblk = lib.passl_b200_bn_reduce_blocks(P, C)
    part = torch.empty((nblk, 2, C), dtype=torch.float32, device=y.device)
    relu = int(relu)
    if relu and mask_bits is not None:
        relu, z = 3, mask_bits          # 1-bit mask from bn_apply_mask: the activation tensor is never read
    elif relu and not want_dres:
        # no residual entered the ReLU: z = relu(fma(y, scale, shift)) — recompute the mask from y, never read z (msss rows 2, 3)
        relu, z = 2, msss[2:4]
    _lib.check(lib.passl_b200_bn_bwd_reduce(_ptr(y), _ptr(dz), _ptr(z), _ptr(msss[0]), _ptr(msss[1]), _ptr(part), P, C,
                                            int(relu), _stream()), "bn_bwd_reduce")
    sums = torch.empty((2, C), dtype=torch.float32, device=y.device)
    coef = torch.empty((3, C), dtype=torch.float32, device=y.device)
    _lib.check(lib.passl_b200_bn_bwd_finalize(_ptr(part), nblk, _ptr(sums), _ptr(dgamma), _ptr(dbeta), _ptr(gamma), _ptr(msss[0]),
                                              _ptr(msss[1]), P, _ptr(coef), C, _stream()), "bn_bwd_finalize")
    dy = torch.empty_like(y)
    dres = torch.empty_like(y) if want_dres else None
    _lib.check(lib.passl_b200_bn_bwd_apply(_ptr(y), _ptr(dz), _ptr(z), _ptr(coef), _ptr(dy), _ptr(dres), P, C, int(relu), _stream()),
               "bn_bwd_apply")
    return dy, dres, sums


def im2col_nchw(x, R, S, stride, pad, kpad):
    _need_cuda(x)
    lib = _lib.load()
    N, C, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W + 2 * pad - S) // stride + 1
    out = torch.empty((N * Ho * Wo, kpad), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.passl_b200_im2col_nchw_f32(_ptr(x), _ptr(out), N, C, H, W, R, S, stride, pad, kpad, _stream()), "im2col")
    return out, Ho, Wo


def maxpool_fwd(x):
    lib = _lib.load()
    N, H, W, C = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty((N, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
    arg = torch.empty((N, Ho, Wo, C), dtype=torch.int8, device=x.device)
    _lib.check(lib.passl_b200_maxpool3x3s2_fwd(_ptr(x), _ptr(y), _ptr(arg), N, H, W, C, _stream()), "maxpool_fwd")
    return y, arg


def bn_relu_maxpool_fwd(y, msss):
    """maxpool3x3/2(relu(y*scale + shift)) for the stem: NHWC bf16 y -> (pooled, argmax int8)."""
    lib = _lib.load()
    N, H, W, C = y.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out = torch.empty((N, Ho, Wo, C), dtype=torch.bfloat16, device=y.device)
    arg = torch.empty((N, Ho, Wo, C), dtype=torch.int8, device=y.device)
    _lib.check(lib.passl_b200_bn_relu_maxpool3x3s2_fwd(_ptr(y), _ptr(msss[2]), _ptr(msss[3]), _ptr(out), _ptr(arg), N, H, W, C,
                                                       _stream()), "bn_relu_maxpool3x3s2_fwd")
    return out, arg


def maxpool_bwd(dy, arg, x_shape):
    lib = _lib.load()
    N, H, W, C = x_shape
    dx = torch.empty(x_shape, dtype=torch.bfloat16, device=dy.device)
    _lib.check(lib.passl_b200_maxpool3x3s2_bwd(_ptr(dy), _ptr(arg), _ptr(dx), N, H, W, C, _stream()), "maxpool_bwd")
    return dx


def avgpool_fwd(x, want_f32=False):
    lib = _lib.load()
    N, H, W, C = x.shape
    y = torch.empty((N, C), dtype=torch.bfloat16, device=x.device)
    yf = torch.empty((N, C), dtype=torch.float32, device=x.device) if want_f32 else None
    _lib.check(lib.passl_b200_avgpool_fwd(_ptr(x), _ptr(y), _ptr(yf), N, H * W, C, _stream()), "avgpool_fwd")
    return y, yf


def avgpool_bwd(dy, x_shape):
    lib = _lib.load()
    N, H, W, C = x_shape
    dx = torch.empty(x_shape, dtype=torch.bfloat16, device=dy.device)
    _lib.check(lib.passl_b200_avgpool_bwd(_ptr(dy.contiguous()), _ptr(dx), N, H * W, C, _stream()), "avgpool_bwd")
    return dx


def ntxent_co2_fwd(S, n, m, rank, co2_weight=3.0):
    """S fp32 [2n, 2m] -> (out fp32[4] = loss, acc1, contrast, co2; stats workspace for the backward)."""
    lib = _lib.load()
    assert S.dtype == torch.float32 and S.is_contiguous() and S.shape == (2 * n, 2 * m)
    out = torch.empty(4, dtype=torch.float32, device=S.device)
    nb = lib.passl_b200_ntxent_workspace_bytes(n)
    ws = torch.empty(nb, dtype=torch.uint8, device=S.device)
    _lib.check(lib.passl_b200_ntxent_co2_fwd(_ptr(S), n, m, rank, float(co2_weight), _ptr(out), _ptr(ws), nb, _stream()),
               "ntxent_co2_fwd")
    return out, ws


def ntxent_co2_bwd(S, ws, n, m, rank, co2_weight=3.0, dloss=None):
    lib = _lib.load()
    dS = torch.empty(S.shape, dtype=torch.bfloat16, device=S.device)
    _lib.check(lib.passl_b200_ntxent_co2_bwd(_ptr(S), n, m, rank, float(co2_weight), _ptr(dloss), _ptr(dS), _ptr(ws),
                                             _stream()), "ntxent_co2_bwd")
    return dS


def cast_f32(x_bf16, out=None):
    lib = _lib.load()
    if out is None:
        out = torch.empty(x_bf16.shape, dtype=torch.float32, device=x_bf16.device)
    _lib.check(lib.passl_b200_cast_bf16_to_f32(_ptr(x_bf16), _ptr(out), x_bf16.numel(), _stream()), "cast_bf16_to_f32")
    return out


def bn_global_affine(running_mean, running_var, gamma, beta, eps=1e-5):
    lib = _lib.load()
    C = running_mean.numel()
    out = torch.empty((4, C), dtype=torch.float32, device=running_mean.device)
    _lib.check(lib.passl_b200_bn_global_affine(_ptr(running_mean), _ptr(running_var), _ptr(gamma), _ptr(beta), _ptr(out[0]),
                                               _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), float(eps), C, _stream()),
               "bn_global_affine")
    return out


def axpy(y, x, a=1.0):
    """y += a * x (fp32, in place)."""
    lib = _lib.load()
    assert y.dtype == torch.float32 and x.dtype == torch.float32 and y.numel() == x.numel()
    assert y.is_contiguous() and x.is_contiguous()
    _lib.check(lib.passl_b200_axpy_f32(_ptr(y), _ptr(x), float(a), y.numel(), _stream()), "axpy_f32")
    return y
