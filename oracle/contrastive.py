"""CPU oracle (TEST INFRASTRUCTURE ONLY — never imported by the product path) for the contrastive heads.

Each function restates, line by line, the reference math it cites (PaddlePaddle/PASSL @ 5c7359b).  Paddle itself is not
installable here, so Paddle op semantics are restated explicitly in numpy float64:
  * nn.CrossEntropyLoss()            = mean_i( logsumexp(logits_i) - logits_i[label_i] )
  * softmax_with_cross_entropy(soft) = per-row  -(sum_j label_ij * log_softmax(logits)_ij)   shape [n,1]
  * kl_div(input=log p, label=q, 'batchmean') = sum_ij q_ij (log q_ij - log p_ij) / n         (0 where q == 0)
  * F.normalize(x, axis) = x / max(||x||_2, 1e-12);  l2_normalize = x / sqrt(sum x^2 + 1e-12)
Parity status: pinned against the reference's own Python source executed over a torch-backed `paddle` shim
(tests/golden/make_golden.py -> tests/golden/*.npz); Paddle's kernels themselves cannot run here.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------------------
# helpers (Paddle op semantics)
# ------------------------------------------------------------------------------------------------------------
def logsumexp(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    return (m + np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True))).squeeze(axis)


def log_softmax(x, axis=-1):
    return x - np.expand_dims(logsumexp(x, axis), axis)


def cross_entropy_mean(logits, labels):
    """paddle.nn.CrossEntropyLoss() with hard int64 labels, reduction='mean'."""
    lsm = log_softmax(logits.astype(np.float64), -1)
    return -np.mean(lsm[np.arange(logits.shape[0]), labels])


def f_normalize(x, axis=1, eps=1e-12):
    """paddle.nn.functional.normalize: x / max(||x||, eps)."""
    n = np.sqrt(np.sum(x.astype(np.float64) ** 2, axis=axis, keepdims=True))
    return x / np.maximum(n, eps)


def l2_normalize(x, axis=-1, eps=1e-12):
    """passl/nn/norm.py:18-40 / fluid.layers.l2_normalize: x / sqrt(sum x^2 + eps)."""
    return x / np.sqrt(np.sum(x.astype(np.float64) ** 2, axis=axis, keepdims=True) + eps)


def topk_accuracy(logits, labels, topk=(1, 5)):
    """contrastive_head.py:63-78 `accuracy`: percentage of rows whose label is among the k largest logits."""
    maxk = max(topk)
    pred = np.argsort(-logits, axis=1, kind="stable")[:, :maxk]
    correct = pred == labels.reshape(-1, 1)
    return [correct[:, :k].sum() * 100.0 / logits.shape[0] for k in topk]


# ------------------------------------------------------------------------------------------------------------
# MoCo v1/v2 : moco.py:154-185 + contrastive_head.py:37-60
# ------------------------------------------------------------------------------------------------------------
def moco_logits(q, k, queue_dk):
    """q,k [N,C] (already normalised); queue_dk [C,K] as the reference stores it.  Returns (l_pos [N,1], l_neg [N,K])."""
    l_pos = np.sum(q * k, axis=1)[:, None]          # moco.py:178
    l_neg = q @ queue_dk                             # moco.py:180
    return l_pos, l_neg


def contrastive_head(pos, neg, temperature):
    """contrastive_head.py:47-59.  Returns dict(loss, acc1, acc5, logits, labels)."""
    n = pos.shape[0]
    logits = np.concatenate((pos, neg), axis=1).astype(np.float64)
    logits = logits / temperature
    labels = np.zeros((n,), dtype=np.int64)
    loss = cross_entropy_mean(logits, labels)
    acc1, acc5 = topk_accuracy(logits, labels, (1, 5))
    return dict(loss=loss, acc1=acc1, acc5=acc5, logits=logits, labels=labels)


def moco_infonce_grad_q(q, k, queue_dk, temperature):
    """d loss / d q for the MoCo head (k and queue carry no gradient: moco.py:162-180)."""
    l_pos, l_neg = moco_logits(q.astype(np.float64), k.astype(np.float64), queue_dk.astype(np.float64))
    logits = np.concatenate((l_pos, l_neg), 1) / temperature
    p = np.exp(log_softmax(logits, -1))
    p[:, 0] -= 1.0
    g = p / (q.shape[0] * temperature)              # d mean-CE / d (unscaled logit)
    return g[:, :1] * k + g[:, 1:] @ queue_dk.T


def dequeue_and_enqueue(queue_dk, queue_ptr, keys_all):
    """moco.py:92-105 on numpy arrays (keys_all already gathered).  Returns (queue, ptr); asserts K % batch == 0."""
    K = queue_dk.shape[1]
    batch_size = keys_all.shape[0]
    ptr = int(queue_ptr)
    assert K % batch_size == 0
    queue_dk = queue_dk.copy()
    queue_dk[:, ptr:ptr + batch_size] = keys_all.T
    ptr = (ptr + batch_size) % K
    return queue_dk, np.int64(ptr)


def momentum_update(param_k, param_q, m):
    """moco.py:82-90: param_k * m + param_q * (1 - m)."""
    return param_k * m + param_q * (1.0 - m)


# ------------------------------------------------------------------------------------------------------------
# MoCo v3 : mocov3.py:187-198
# ------------------------------------------------------------------------------------------------------------
def mocov3_contrastive_loss(q, k_all, T, rank=0):
    """q [N,C] local queries, k_all [world*N, C] gathered keys (both un-normalised inputs)."""
    q = f_normalize(q.astype(np.float64), 1)
    k_all = f_normalize(k_all.astype(np.float64), 1)
    logits = np.einsum("nc,mc->nm", q, k_all) / T
    N = logits.shape[0]
    labels = np.arange(N, dtype=np.int64) + N * rank
    return cross_entropy_mean(logits, labels) * (2 * T), logits, labels


# ------------------------------------------------------------------------------------------------------------
# CLIP : clip.py:320-335 + clip_head.py:27-35
# ------------------------------------------------------------------------------------------------------------
def clip_logits(image_features, text_features, logit_scale_log):
    i = image_features.astype(np.float64)
    t = text_features.astype(np.float64)
    i = i / np.linalg.norm(i, axis=-1, keepdims=True)
    t = t / np.linalg.norm(t, axis=-1, keepdims=True)
    s = np.exp(np.float64(logit_scale_log))
    return (s * i) @ t.T, (s * t) @ i.T


def clip_head(img_logits, text_logits, img_labels, text_labels):
    il = cross_entropy_mean(img_logits, img_labels)
    tl = cross_entropy_mean(text_logits, text_labels)
    return dict(img_loss=il, text_loss=tl, loss=il + tl)


# ------------------------------------------------------------------------------------------------------------
# SimCLR NT-Xent + CO2 : simclr_contrastive_head.py:42-102
# ------------------------------------------------------------------------------------------------------------
LARGE_NUM = 1e9


def simclr_head(hidden1, hidden2, temperature):
    h1 = hidden1.astype(np.float64)
    h2 = hidden2.astype(np.float64)
    n = h1.shape[0]
    labels = np.eye(n, 2 * n)                       # one_hot(arange(n), 2n)
    masks = np.eye(n)
    logits_aa = h1 @ h1.T / temperature - masks * LARGE_NUM
    logits_bb = h2 @ h2.T / temperature - masks * LARGE_NUM
    logits_ab = h1 @ h2.T / temperature
    logits_ba = h2 @ h1.T / temperature
    loss_a = -np.sum(labels * log_softmax(np.concatenate([logits_ab, logits_aa], 1)), 1, keepdims=True)
    loss_b = -np.sum(labels * log_softmax(np.concatenate([logits_ba, logits_bb], 1)), 1, keepdims=True)
    contrast_loss = loss_a + loss_b
    logit_a = np.concatenate([logits_aa, logits_ab - masks * LARGE_NUM], 1)
    logit_b = np.concatenate([logits_ba - masks * LARGE_NUM, logits_bb], 1)
    log_a, log_b = log_softmax(logit_a), log_softmax(logit_b)
    a, b = np.exp(log_a), np.exp(log_b)

    def kl_batchmean(log_p, q):
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(q > 0, q * (np.log(q) - log_p), 0.0)
        return t.sum() / log_p.shape[0]

    co2 = kl_batchmean(log_a, b) + kl_batchmean(log_b, a)
    total = contrast_loss + 3 * co2
    loss = total.mean()
    acc1 = np.mean(np.argmax(logits_ab, 1) == np.arange(n))   # layers.accuracy returns a fraction
    return dict(loss=loss, acc1=acc1, contrast=contrast_loss.mean(), co2=co2)
