"""Fused InfoNCE (fp32 SIMT variant) + embedding utilities vs the CPU oracle (oracle/contrastive.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _moco_inputs(N, D, K, seed):
    rng = np.random.RandomState(seed)
    q = rng.randn(N, D).astype(np.float32)
    k = rng.randn(N, D).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    k /= np.linalg.norm(k, axis=1, keepdims=True)
    k = (0.5 * q + 0.5 * k).astype(np.float32)
    queue = rng.randn(D, K).astype(np.float32)
    queue /= np.linalg.norm(queue, axis=0, keepdims=True)
    return q, k, queue


@pytest.mark.parametrize("N,D,K,T", [(16, 128, 65536, 0.2), (256, 128, 8192, 0.07), (37, 256, 1000, 0.2)])
def test_moco_infonce_fp32_matches_oracle(N, D, K, T):
    from oracle import contrastive as O
    from passl_b200 import kernels as K_
    q, k, queue = _moco_inputs(N, D, K, 1234)
    l_pos, l_neg = O.moco_logits(q.astype(np.float64), k.astype(np.float64), queue.astype(np.float64))
    ref = O.contrastive_head(l_pos, l_neg, T)
    gref = O.moco_infonce_grad_q(q, k, queue, T)
    qd, kd = torch.from_numpy(q).cuda(), torch.from_numpy(k).cuda()
    queue_kd = torch.from_numpy(np.ascontiguousarray(queue.T)).cuda()
    out, lse, tgt, _ = K_.simce_fwd(qd, queue_kd, pos=kd, scale=1.0 / T)
    dq = K_.simce_bwd(qd, queue_kd, lse, tgt, pos=kd, scale=1.0 / T)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    assert abs(out[0] - ref["loss"]) <= 1e-3 * abs(ref["loss"]), (out[0], ref["loss"])   # 1e-3 rel fp32 (BASELINE.json)
    assert out[1] == pytest.approx(ref["acc1"], abs=1e-3) and out[2] == pytest.approx(ref["acc5"], abs=1e-3)
    # logits parity through the saved target logit and LSE
    lse_ref = O.logsumexp(ref["logits"], -1)
    np.testing.assert_allclose(lse.cpu().numpy(), lse_ref, rtol=1e-3)
    np.testing.assert_allclose(tgt.cpu().numpy(), ref["logits"][:, 0], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(dq.cpu().numpy(), gref, rtol=1e-3, atol=1e-3 * np.abs(gref).max())


def test_label_mode_mocov3_and_bf16_keys():
    from oracle import contrastive as O
    from passl_b200 import kernels as K_
    rng = np.random.RandomState(7)
    N, D, world, T, rank = 48, 256, 4, 0.2, 2
    q = rng.randn(N, D).astype(np.float32)
    k_all = rng.randn(world * N, D).astype(np.float32)
    loss_ref, logits, labels = O.mocov3_contrastive_loss(q, k_all, T, rank)
    qn = torch.from_numpy(O.f_normalize(q).astype(np.float32)).cuda()
    kn = torch.from_numpy(O.f_normalize(k_all).astype(np.float32)).cuda()
    lab = torch.from_numpy(labels).cuda()
    out, lse, tgt, _ = K_.simce_fwd(qn, kn, label=lab, scale=1.0 / T, loss_scale=2 * T)
    torch.cuda.synchronize()
    assert abs(out[0].item() - loss_ref) <= 1e-3 * abs(loss_ref)
    # labels are bit-exact integers: arange(N) + N*rank
    assert lab.dtype == torch.int64 and torch.equal(lab.cpu(), torch.arange(N) + N * rank)
    out_b, _, _, _ = K_.simce_fwd(qn, kn.bfloat16(), label=lab, scale=1.0 / T, loss_scale=2 * T)
    assert abs(out_b[0].item() - loss_ref) <= 1e-2 * abs(loss_ref)      # 1e-2 rel with bf16 keys


def test_queue_ring_buffer_bit_exact():
    from oracle import contrastive as O
    from passl_b200 import kernels as K_
    rng = np.random.RandomState(3)
    D, K, Bg = 128, 4096, 256
    queue = rng.randn(D, K).astype(np.float32)
    ptr = np.int64(K - 2 * Bg)
    qd = torch.from_numpy(np.ascontiguousarray(queue.T)).cuda()
    qb = qd.bfloat16()
    pd = torch.tensor([int(ptr)], dtype=torch.int64, device="cuda")
    for step in range(5):   # wraps around
        keys = rng.randn(Bg, D).astype(np.float32)
        queue, ptr = O.dequeue_and_enqueue(queue, ptr, keys)
        K_.queue_enqueue(torch.from_numpy(keys).cuda(), pd, queue_f32=qd, queue_bf16=qb)
        torch.cuda.synchronize()
        assert int(pd.item()) == int(ptr)
        assert np.array_equal(qd.cpu().numpy(), queue.T)            # bit-exact copy
        assert torch.equal(qb.cpu(), torch.from_numpy(np.ascontiguousarray(queue.T)).bfloat16())
    with pytest.raises(AssertionError):
        K_.queue_enqueue(torch.zeros(100, D, device="cuda"), pd, queue_f32=qd)


@pytest.mark.parametrize("mode", ["normalize", "l2_normalize", "clip"])
def test_l2norm_fwd_bwd(mode):
    from oracle import contrastive as O
    from passl_b200 import kernels as K_
    torch.manual_seed(0)
    x = torch.randn(300, 128, device="cuda")
    y, yb, inv = K_.l2norm_fwd(x, mode=mode, want_bf16=True)
    xr = x.double().cpu().requires_grad_(True)
    if mode == "normalize":
        yr = xr / xr.norm(dim=1, keepdim=True).clamp_min(1e-12)
        ref_np = O.f_normalize(x.cpu().numpy().astype(np.float64))
    elif mode == "l2_normalize":
        yr = xr / (xr.pow(2).sum(1, keepdim=True) + 1e-12).sqrt()
        ref_np = O.l2_normalize(x.cpu().numpy().astype(np.float64))
    else:
        yr = xr / xr.norm(dim=1, keepdim=True)
        ref_np = x.cpu().numpy().astype(np.float64)
        ref_np = ref_np / np.linalg.norm(ref_np, axis=-1, keepdims=True)
    np.testing.assert_allclose(y.cpu().numpy(), ref_np, rtol=1e-5, atol=1e-6)
    g = torch.randn(300, 128, device="cuda")
    yr.backward(g.double().cpu())
    dx, _ = K_.l2norm_bwd(g, y, inv, mode=mode)
    torch.cuda.synchronize()
    np.testing.assert_allclose(dx.cpu().numpy(), xr.grad.numpy(), rtol=1e-4, atol=1e-5)
    assert torch.equal(yb, y.bfloat16())


def test_ema_update():
    from oracle import contrastive as O
    from passl_b200 import kernels as K_
    torch.manual_seed(0)
    n = 1_000_003
    k = torch.randn(n, device="cuda")
    q = torch.randn(n, device="cuda")
    kb = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    ref = O.momentum_update(k.cpu().numpy(), q.cpu().numpy(), np.float32(0.999))
    K_.ema_update(k, q, 0.999, k_bf16=kb)
    torch.cuda.synchronize()
    np.testing.assert_allclose(k.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)
    assert torch.equal(kb, k.bfloat16())
