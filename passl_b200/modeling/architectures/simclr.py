"""SimCLR (passl_v110/modeling/architectures/simclr.py:27-80): both views are concatenated through ONE encoder pass
(shared BatchNorm statistics over 2N images), l2-normalised, split, and fed to the head."""
import torch
import torch.nn as nn

from ...loss.contrastive import l2_normalize
from ..registry import MODELS, build_backbone, build_neck, build_head


class _ConcatImages(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        out = torch.empty((a.shape[0] + b.shape[0],) + tuple(a.shape[1:]), dtype=a.dtype, device=a.device)
        out[:a.shape[0]].copy_(a)
        out[a.shape[0]:].copy_(b)
        return out

    @staticmethod
    def backward(ctx, g):
        return None, None


@MODELS.register()
class SimCLR(nn.Module):
    def __init__(self, backbone, neck=None, head=None, dim=128, T=0.5):
        super().__init__()
        self.T = T
        self.encoder = nn.Sequential(build_backbone(backbone), build_neck(neck))
        self.backbone = self.encoder[0]
        self.head = build_head(head)

    def train_iter(self, *inputs, **kwargs):
        img_q, img_k = inputs
        img_con = _ConcatImages.apply(img_q, img_k)       # paddle.concat([img_q, img_k])
        con = self.encoder(img_con)
        con = l2_normalize(con)                            # layers.l2_normalize(con, -1)
        # layers.split(con, 2, dim=0) + head(q, k): the fused head consumes [q; k] directly
        return self.head.forward_fused(con, img_q.shape[0])

    def forward(self, *inputs, mode='train', **kwargs):
        if mode == 'train':
            return self.train_iter(*inputs, **kwargs)
        elif mode == 'extract':
            return self.backbone(*inputs)
        else:
            raise Exception("No such mode: {}".format(mode))
