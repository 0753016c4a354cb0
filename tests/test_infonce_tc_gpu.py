"""tcgen05 fused InfoNCE forward vs the CPU oracle (fp64) and vs the fp32 SIMT variant.

Tolerances (BASELINE.json north_star): loss / logits within 1e-2 relative for the bf16 path; labels and rank
counters are integer-exact given identical logits.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(N, D, K, seed):
    rng = np.random.RandomState(seed)
    q = rng.randn(N, D).astype(np.float32)
    k = rng.randn(N, D).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    k /= np.linalg.norm(k, axis=1, keepdims=True)
    k = 0.6 * q + 0.4 * k
    k /= np.linalg.norm(k, axis=1, keepdims=True)
    queue = rng.randn(K, D).astype(np.float32)
    queue /= np.linalg.norm(queue, axis=1, keepdims=True)
    return q, k.astype(np.float32), queue


@pytest.mark.parametrize("N,D,K,T", [(256, 128, 65536, 0.2), (16, 128, 65536, 0.2), (128, 128, 4096, 0.07),
                                     (200, 128, 1000, 0.2), (96, 256, 2048, 0.2), (64, 64, 640, 0.1)])
def test_moco_form_matches_oracle(N, D, K, T):
    from oracle import contrastive as O
    from passl_b200 import kernels as K_
    q, k, queue = _inputs(N, D, K, 1234)
    qb = torch.from_numpy(q).cuda().bfloat16()
    kb = torch.from_numpy(queue).cuda().bfloat16()
    kd = torch.from_numpy(k).cuda()
    # oracle on the bf16-rounded operands (isolates kernel error from the quantisation error) ...
    q64 = qb.float().cpu().numpy().astype(np.float64)
    queue64 = kb.float().cpu().numpy().astype(np.float64)
    l_pos, l_neg = O.moco_logits(q64, k.astype(np.float64), queue64.T)
    ref = O.contrastive_head(l_pos, l_neg, T)
    # ... and on the original fp32 operands (the end-to-end bf16 tolerance of BASELINE.json)
    l_pos0, l_neg0 = O.moco_logits(q.astype(np.float64), k.astype(np.float64), queue.astype(np.float64).T)
    ref0 = O.contrastive_head(l_pos0, l_neg0, T)

    out, lse, tgt, rows = K_.infonce_tc_fwd(qb, kb, pos=kd, scale=1.0 / T, want_rows=True)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    lse_ref = O.logsumexp(ref["logits"], -1)
    np.testing.assert_allclose(lse.cpu().numpy(), lse_ref, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(tgt.cpu().numpy(), ref["logits"][:, 0], rtol=1e-5, atol=1e-5)
    assert abs(out[0] - ref["loss"]) <= 1e-4 * abs(ref["loss"]), (out[0], ref["loss"])
    assert abs(out[0] - ref0["loss"]) <= 1e-2 * abs(ref0["loss"]), (out[0], ref0["loss"])
    assert out[1] == pytest.approx(ref["acc1"], abs=100.0 / N + 1e-3)
    assert out[2] == pytest.approx(ref["acc5"], abs=100.0 / N + 1e-3)


def test_label_mode_and_exclusion_match_simt_variant():
    from passl_b200 import kernels as K_
    torch.manual_seed(5)
    N, D, K = 300, 128, 1536
    q = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=1)
    keys = torch.nn.functional.normalize(torch.randn(K, D, device="cuda"), dim=1)
    lab = (torch.arange(N, device="cuda") + 600).to(torch.int64)
    excl = (torch.arange(N, device="cuda") + 100).to(torch.int32)
    qb, kb = q.bfloat16(), keys.bfloat16()
    o_tc, lse_tc, tgt_tc, _ = K_.infonce_tc_fwd(qb, kb, label=lab, excl=excl, scale=10.0, loss_scale=0.4)
    o_f, lse_f, tgt_f, _ = K_.simce_fwd(qb.float(), kb, label=lab, excl=excl, scale=10.0, loss_scale=0.4)
    torch.cuda.synchronize()
    assert torch.allclose(lse_tc, lse_f, rtol=1e-4, atol=1e-4), (lse_tc - lse_f).abs().max()
    assert torch.allclose(tgt_tc, tgt_f, rtol=1e-4, atol=1e-4)
    assert torch.allclose(o_tc, o_f, rtol=1e-4, atol=1e-2), (o_tc, o_f)


def test_known_answers():
    """SURVEY.md §8c golden identities: zero embeddings -> ln(K+1); q==k, queue orthogonal -> ln(1+K e^{-1/T})."""
    from passl_b200 import kernels as K_
    N, D, K, T = 128, 128, 65536, 0.2
    z = torch.zeros(N, D, device="cuda")
    out, _, _, _ = K_.infonce_tc_fwd(z.bfloat16(), torch.zeros(K, D, device="cuda").bfloat16(), pos=z, scale=1 / T)
    assert abs(out[0].item() - np.log(K + 1)) < 1e-4          # 11.09035
    q = torch.zeros(N, D, device="cuda")
    q[:, 0] = 1.0
    queue = torch.zeros(K, D, device="cuda")
    queue[:, 1] = 1.0
    out, _, _, _ = K_.infonce_tc_fwd(q.bfloat16(), queue.bfloat16(), pos=q, scale=1 / T)
    assert abs(out[0].item() - np.log(1 + K * np.exp(-1 / T))) < 1e-3
    assert out[1].item() == 100.0 and out[2].item() == 100.0
