"""CPU restatement (TEST INFRASTRUCTURE / bench.py cpu_baseline + --impl reference ONLY) of one full MoCo v1/v2 training
iteration of the reference:
  passl_v110/modeling/architectures/moco.py:154-185 (train_iter: q = normalize(encoder_q(img_q)); momentum update of the key
      encoder :82-90; k = normalize(encoder_k(img_k)) under no_grad — its BatchNorm layers use the GLOBAL statistics,
      freeze_batchnorm_statictis moco.py:74 / modules/freeze.py:17-23; l_pos / l_neg :178-180; head; _dequeue_and_enqueue :92-105)
  heads/contrastive_head.py:37-60 ([l_pos | l_neg] / T, labels 0, CrossEntropyLoss)
  backbones resnetimagenet.py:93-246, necks/base_neck.py:67-94 (NonLinearNeckV1), configs/moco/moco_v2_r50.yaml
  optimizer paddle.optimizer.Momentum with L2Decay folded into the gradient (solver/optimizer.py:24; oracle/optim.py)
in torch-CPU (float32 for timing, float64 / quantisation-matched for parity tests) with autograd.  Paddle itself cannot be
installed here; parity status of the pieces: oracle/resnet.py, oracle/contrastive.py (pinned), oracle/optim.py
(update rule unpinned)."""
import math

import torch
import torch.nn.functional as F

from . import resnet as R
from .simclr_step import init_params as _init_r50


def init_params(seed=0, dtype=torch.float32, K=65536, dim=128, layers=(3, 4, 6, 3)):
    """encoder_q parameters (ResNet-50 kaiming fan_out + NonLinearNeckV1), a copy for encoder_k, running statistics of the key
    encoder's BatchNorm layers (mean 0 / variance 1 as constructed), and the queue."""
    g = torch.Generator().manual_seed(seed)
    pq = {k: v.detach() for k, v in _init_r50(seed, dtype, layers).items() if not k.startswith("neck.")}
    for name, (cin, cout) in (("neck.fc1", (2048, 2048)), ("neck.fc2", (2048, dim))):
        pq[name + ".weight"] = torch.randn(cout, cin, generator=g, dtype=dtype) * math.sqrt(2.0 / cin)
        pq[name + ".bias"] = torch.zeros(cout, dtype=dtype)
    for v in pq.values():
        v.requires_grad_(True)
    pk = {k: v.detach().clone() for k, v in pq.items()}
    for k in list(pk):
        if k.endswith(".bn.weight"):
            c = pk[k].shape[0]
            pk[k[:-len("weight")] + "_mean"] = torch.zeros(c, dtype=dtype)
            pk[k[:-len("weight")] + "_variance"] = torch.ones(c, dtype=dtype)
    queue = F.normalize(torch.randn(dim, K, generator=g, dtype=dtype), dim=0)        # moco.py:77-78, layout [dim, K]
    return dict(q=pq, k=pk, queue=queue, ptr=0, velocity={})


def encode(img, p, ugs=False, q=False):
    feat = R.resnet_forward(img, p, ugs=ugs, q=q)
    return R.neck_v1(feat, p, prefix="neck.", q=q)


def train_step(state, img_q, img_k, lr=0.03, T=0.2, m=0.999, momentum=0.9, wd=1e-4, q=False, update_bn_stats=None):
    """One iteration; returns dict(loss, acc1, acc5).  q: quantisation-matched mode (oracle/resnet.py).
    update_bn_stats(state, stats) lets a parity test feed the query encoder's batch statistics to wherever it tracks them."""
    pq, pk = state["q"], state["k"]
    emb_q = encode(img_q, pq, q=q)
    qn = F.normalize(emb_q, dim=1)                                               # moco.py:159
    with torch.no_grad():
        for name, w in pq.items():                                               # moco.py:82-90
            pk[name].mul_(m).add_(w.detach() * (1.0 - m))
        kn = F.normalize(encode(img_k, pk, ugs=True, q=q), dim=1)                # moco.py:171-172 (global-statistics BN)
    if q:   # the fused loss kernel multiplies bf16 queries with the bf16 queue mirror; the positive pair uses fp32 keys
        qb = R.Qf(qn)
        l_pos = (qb * kn).sum(1, keepdim=True)
        l_neg = qb @ R._round_bf16(state["queue"])
    else:
        l_pos = (qn * kn).sum(1, keepdim=True)                                   # moco.py:178
        l_neg = qn @ state["queue"]                                              # moco.py:180
    logits = torch.cat([l_pos, l_neg], 1) / T                                    # contrastive_head.py:47-49
    labels = torch.zeros(logits.shape[0], dtype=torch.long)
    loss = F.cross_entropy(logits, labels)
    with torch.no_grad():
        top5 = logits.topk(5, dim=1).indices
        acc1 = (top5[:, 0] == 0).double().mean().item() * 100
        acc5 = (top5 == 0).any(1).double().mean().item() * 100
    grads = torch.autograd.grad(loss, list(pq.values()))
    with torch.no_grad():
        B = kn.shape[0]                                                          # moco.py:92-105
        assert state["queue"].shape[1] % B == 0
        ptr = state["ptr"]
        state["queue"][:, ptr:ptr + B] = kn.t()
        state["ptr"] = (ptr + B) % state["queue"].shape[1]
        for (name, w), g in zip(pq.items(), grads):                              # oracle/optim.py::momentum
            v = state["velocity"].setdefault(name, torch.zeros_like(w))
            v.mul_(momentum).add_(g + wd * w)
            w.sub_(lr * v)
    return dict(loss=float(loss.detach()), acc1=acc1, acc5=acc5)


def infonce_head_unfused(q, k, queue, T=0.2):
    """The reference's unfused head alone (matmul -> concat -> /T -> CE), for the cpu_baseline of the fused-InfoNCE metric."""
    l_pos = (q * k).sum(1, keepdim=True)
    l_neg = q @ queue
    logits = torch.cat([l_pos, l_neg], 1) / T
    return F.cross_entropy(logits, torch.zeros(q.shape[0], dtype=torch.long))
