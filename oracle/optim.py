"""CPU restatement (fp64, one tensor at a time) of the parameter-update rules on the path — TEST INFRASTRUCTURE, not a product
path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

The arithmetic lives in PaddlePaddle's C++ operators, a third-party dependency that is not in /root/reference (docs/INSTALL.md:36,52
installs `paddlepaddle-gpu`, 2.1.1.post101 in the pinned example; the v2.5 tree needs 2.4+).  Their published update rules are
restated here and anchored on the reference's call sites:

  momentum(p, g, v)          paddle.optimizer.Momentum registered at passl_v110/solver/optimizer.py:24, used by
                             configs/moco/moco_v2_r50.yaml:89-92 with `weight_decay` = L2Decay folded into the gradient
  lars_momentum(p, g, v)     paddle.fluid.optimizer.LarsMomentum registered at passl_v110/solver/optimizer.py:25, used by
                             configs/simclr/simclr_r50_IM.yaml:116-120 (lars_coeff 0.001, lars_weight_decay, exclude list)
  adamw(p, g, m, v)          _C_ops.adamw called by passl/optimizer/adamw.py:101-137 with the attributes epsilon, beta1, beta2,
                             with_decay, coeff = weight_decay, lr_ratio = 1.0, and beta^step passed in as beta1_pow / beta2_pow

PARITY UNPINNED for the three C++ rules: no golden vector can be produced without PaddlePaddle; the pin is the published formula
plus the call sites above.  (The v2.5 tree's own Python LARS, passl/optimizer/momentum_lars.py:96-114, is a different rule —
trust ratio on ||g + wd p|| — and is not what the SimCLR recipe of the v110 tree runs; it is not built.)
"""
import torch


def momentum(p, g, v, lr, mu, wd):
    """velocity = mu * velocity + (g + wd * p);  p -= lr * velocity   (use_nesterov = False, L2Decay regulariser)."""
    g = g + wd * p
    v = mu * v + g
    return p - lr * v, v


def lars_momentum(p, g, v, lr, mu, wd, coeff=0.001, eps=0.0):
    """local_lr = lr * coeff * ||p|| / (||g|| + wd ||p|| + eps) when wd, ||p||, ||g|| > 0, else lr;
    velocity = mu * velocity + local_lr * (g + wd * p);  p -= velocity."""
    pn, gn = p.norm(), g.norm()
    local_lr = lr * coeff * pn / (gn + wd * pn + eps) if (wd > 0 and pn > 0 and gn > 0) else lr
    v = mu * v + local_lr * (g + wd * p)
    return p - v, v


def adamw(p, g, m, v, lr, beta1, beta2, eps, wd, step, lr_ratio=1.0):
    """p *= 1 - lr * coeff (with_decay);  m, v moments;  p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)."""
    lr = lr * lr_ratio
    p = p * (1.0 - lr * wd)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    denom = v.sqrt() / (1 - beta2 ** step) ** 0.5 + eps
    return p - lr / (1 - beta1 ** step) * m / denom, m, v
