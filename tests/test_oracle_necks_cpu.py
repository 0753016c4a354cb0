"""oracle/resnet.py necks against golden vectors produced by the reference's own neck classes
(tests/golden/make_golden_models.py: passl_v110/modeling/necks/base_neck.py run over the paddle shim)."""
import os

import numpy as np
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_necks.npz"))


def t(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float64))


def test_nonlinear_neck_v1():
    import oracle.resnet as OR
    p = {"fc1.weight": t(G["v1_mlp.0.weight"]).t(), "fc1.bias": t(G["v1_mlp.0.bias"]),       # paddle [in, out] -> [out, in]
         "fc2.weight": t(G["v1_mlp.2.weight"]).t(), "fc2.bias": t(G["v1_mlp.2.bias"])}
    y = OR.neck_v1(t(G["v1_x"]), p)
    np.testing.assert_allclose(y.numpy(), G["v1_y"], rtol=1e-12, atol=1e-12)


def test_linear_neck():
    x = t(G["lin_x"]).mean(dim=(2, 3))
    y = x @ t(G["lin_fc.weight"]) + t(G["lin_fc.bias"])
    np.testing.assert_allclose(y.numpy(), G["lin_y"], rtol=1e-12, atol=1e-12)
    # the CUDA LinearNeck stores the weight [out, in]: same contraction
    import torch.nn.functional as F
    np.testing.assert_allclose(F.linear(x, t(G["lin_fc.weight"]).t(), t(G["lin_fc.bias"])).numpy(), G["lin_y"], rtol=1e-12, atol=1e-12)


def test_nonlinear_neck_fc3():
    import oracle.resnet as OR
    p = {}
    for i, (fc, bn) in enumerate([(0, 1), (3, 4), (6, 7)], start=1):
        p["fc%d.weight" % i] = t(G["fc3_mlp.%d.weight" % fc]).t()
        p["fc%d.bias" % i] = t(G["fc3_mlp.%d.bias" % fc])
        p["bn%d.bn.weight" % i] = t(G["fc3_mlp.%d.weight" % bn])
        p["bn%d.bn.bias" % i] = t(G["fc3_mlp.%d.bias" % bn])
    y = OR.neck_fc3(t(G["fc3_x"]), p)
    np.testing.assert_allclose(y.numpy(), G["fc3_y"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose((y * y).sum(-1).numpy(), 1.0, rtol=1e-9)          # trailing l2_normalize
