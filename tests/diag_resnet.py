"""Layer-by-layer diagnostic: feed every unit of the CUDA ResNet-50 and of the CPU oracle the SAME (CUDA-produced) input
and report the per-unit relative error, so a broken unit stands out from accumulated bf16 drift."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resnet as O  # noqa: E402
from passl_b200 import kernels as K  # noqa: E402
from passl_b200.modeling import build_backbone  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nchw(x):
    return x.float().cpu().double().permute(0, 3, 1, 2).contiguous()


torch.manual_seed(0)
net = build_backbone(dict(name="ResNet", depth=50)).cuda()
p = O.params_from_cuda_module(net)
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 128
img = torch.randn(B, 3, S, S, device="cuda")
x, _ = net.stem.fwd(img)
ref = O.conv_bn(img.cpu().double().bfloat16().double(), p, "stem", stride=2, pad=3)
ref = torch.nn.functional.max_pool2d(ref, 3, 2, 1)
print("stem rel %.4f" % rel(x.permute(0, 3, 1, 2), ref))
inpl, bi = 64, 0
for i, (planes, n) in enumerate(zip([64, 128, 256, 512], [3, 4, 6, 3])):
    for b in range(n):
        s = (1 if i == 0 else 2) if b == 0 else 1
        ds = b == 0 and (s != 1 or inpl != planes * 4)
        blk = net.blocks[bi]
        y, _ = blk.fwd(x)
        r = O.bottleneck(nchw(x), p, "blocks.%d" % bi, s, ds)
        print("block %2d (in %s, stride %d) rel %.4f  | mean |y| %.3f" % (bi, tuple(x.shape), s, rel(y.permute(0, 3, 1, 2), r), y.float().abs().mean().item()))
        x = y
        inpl = planes * 4
        bi += 1
