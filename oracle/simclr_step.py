"""CPU restatement (TEST INFRASTRUCTURE / bench.py cpu_baseline ONLY) of one full SimCLR training iteration of the reference:
  passl_v110/modeling/architectures/simclr.py:52-61 (concat two views -> encoder -> l2_normalize -> split -> head)
  backbones resnetimagenet.py:93-246 (ResNet-50), necks/base_neck.py:209-237 (NonLinearNeckfc3),
  heads/simclr_contrastive_head.py:42-102 (NT-Xent + 3*CO2), optimizer paddle LarsMomentum (solver/optimizer.py:20-26)
in torch-CPU float32 with autograd, using all host threads torch is given.  It is what `bench.py --impl reference` and the
`cpu_baseline` leg time (Paddle itself cannot be installed here: SURVEY.md §8c)."""
import math

import torch
import torch.nn.functional as F

from . import resnet as R

LARGE_NUM = 1e9


def init_params(seed=0, dtype=torch.float32, layers=(3, 4, 6, 3)):
    g = torch.Generator().manual_seed(seed)
    p = {}

    def conv(name, cout, cin, k):
        p[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g, dtype=dtype) * math.sqrt(2.0 / (cout * k * k))
        p[name + ".bn.weight"] = torch.ones(cout, dtype=dtype)
        p[name + ".bn.bias"] = torch.zeros(cout, dtype=dtype)
    conv("stem", 64, 3, 7)
    inpl, bi = 64, 0
    for i, (planes, n) in enumerate(zip([64, 128, 256, 512], layers)):
        for b in range(n):
            s = (1 if i == 0 else 2) if b == 0 else 1
            pre = "blocks.%d" % bi
            conv(pre + ".conv1", planes, inpl, 1)
            conv(pre + ".conv2", planes, planes, 3)
            conv(pre + ".conv3", planes * 4, planes, 1)
            if b == 0 and (s != 1 or inpl != planes * 4):
                conv(pre + ".downsample", planes * 4, inpl, 1)
            inpl = planes * 4
            bi += 1
    for i, (cin, cout) in enumerate([(2048, 2048), (2048, 2048), (2048, 128)], 1):
        p["neck.fc%d.weight" % i] = torch.randn(cout, cin, generator=g, dtype=dtype) * 0.01
        p["neck.fc%d.bias" % i] = torch.zeros(cout, dtype=dtype)
        p["neck.bn%d.bn.weight" % i] = torch.ones(cout, dtype=dtype)
        p["neck.bn%d.bn.bias" % i] = torch.zeros(cout, dtype=dtype)
    for v in p.values():
        v.requires_grad_(True)
    return p


def simclr_head_loss(h1, h2, T, q=False):
    """simclr_contrastive_head.py:52-94 in torch (autograd).  q: quantisation-matched (bf16 embeddings into the similarity GEMM,
    bf16 gradient of the similarity matrix coming back from the row kernel)."""
    n = h1.shape[0]
    eye = torch.eye(n, dtype=h1.dtype)
    h1, h2 = R.Qf(h1, q), R.Qf(h2, q)
    aa = R.Qb(h1 @ h1.t() / T, q) - eye * LARGE_NUM
    bb = R.Qb(h2 @ h2.t() / T, q) - eye * LARGE_NUM
    ab = R.Qb(h1 @ h2.t() / T, q)
    ba = R.Qb(h2 @ h1.t() / T, q)
    lab = torch.arange(n)
    loss_a = F.cross_entropy(torch.cat([ab, aa], 1), lab, reduction="none")
    loss_b = F.cross_entropy(torch.cat([ba, bb], 1), lab, reduction="none")
    log_a = torch.log_softmax(torch.cat([aa, ab - eye * LARGE_NUM], 1), 1)
    log_b = torch.log_softmax(torch.cat([ba - eye * LARGE_NUM, bb], 1), 1)
    a, b = log_a.exp(), log_b.exp()
    kl1 = torch.where(b > 0, b * (log_b - log_a), torch.zeros_like(b)).sum() / n
    kl2 = torch.where(a > 0, a * (log_a - log_b), torch.zeros_like(a)).sum() / n
    return (loss_a + loss_b).mean() + 3 * (kl1 + kl2)


def train_step(p, velocity, img_q, img_k, lr=0.1, T=0.1, momentum=0.9, lars_wd=1e-4, lars_coeff=0.001, exclude=None, q=False):
    """One iteration: forward, backward, LarsMomentum update (python loop per tensor, like the reference). Returns loss.
    `exclude(name) -> bool` marks tensors without decay / trust ratio; None = every tensor decays, which is what the SimCLR YAML's
    exclude list amounts to in the reference (its substrings match none of Paddle's generated names, see
    passl_b200/optimizer/naming.py)."""
    img = torch.cat([img_q, img_k])                                  # simclr.py:55
    feat = R.resnet_forward(img, p, with_pool=True, q=q)
    con = R.neck_fc3(feat, p, prefix="neck.", q=q)
    con = con / torch.sqrt((con * con).sum(-1, keepdim=True) + 1e-12)  # layers.l2_normalize(con, -1), simclr.py:58
    n = img_q.shape[0]
    loss = simclr_head_loss(con[:n], con[n:], T, q=q)
    grads = torch.autograd.grad(loss, list(p.values()))
    with torch.no_grad():
        for (name, w), g in zip(p.items(), grads):
            wd = 0.0 if (exclude is not None and exclude(name)) else lars_wd
            pn, gn = w.norm(), g.norm()
            local_lr = lr * lars_coeff * pn / (gn + wd * pn) if (wd > 0 and pn > 0 and gn > 0) else lr
            v = velocity.setdefault(name, torch.zeros_like(w))
            v.mul_(momentum).add_(local_lr * (g + wd * w))
            w.sub_(v)
    return float(loss.detach())
