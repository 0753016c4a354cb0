"""v110 registry entries of the MAE pre-training path: `MAE_PRETRAIN` (MODELS, architectures/MAE.py:29-56) wrapping the `MAE`
backbone (BACKBONES, backbones/mae.py:318-564) — the names configs/mae/mae_vit_b_pretrain.yaml uses.  The arithmetic is the v2.5
`MaskedAutoencoderViT` (passl/models/mae.py:37-290; the v110 class is its twin): fused masking / token assembly /
masked-patch MSE kernels, tcgen05 attention + GEMMs.

Reference quirk not reproduced: MAE_PRETRAIN.train_iter hands the whole `inputs` tuple to the backbone (MAE.py:42-45) and returns a
bare (loss, pred, mask) tuple, which the v110 hooks (optimizer_hook.py:30: outputs['loss']) cannot consume; here the first input is
the image batch and the result is {'loss', 'pred', 'mask'}."""
import torch.nn as nn

from ...models.mae import MaskedAutoencoderViT
from ..registry import BACKBONES, MODELS, build_backbone


@BACKBONES.register()
class MAE(MaskedAutoencoderViT):
    """backbones/mae.py:318-335: img_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
    decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4., norm_pix_loss=False (same defaults here)."""


@MODELS.register()
class MAE_PRETRAIN(nn.Module):
    def __init__(self, architecture=None, mask_ratio=0.75):
        super().__init__()
        self.backbone = build_backbone(architecture)
        self.mask_ratio = mask_ratio

    def train_iter(self, *inputs, **kwargs):
        loss, pred, mask = self.backbone(inputs[0], self.mask_ratio)
        return {"loss": loss, "pred": pred, "mask": mask}

    def forward(self, *inputs, mode='train', **kwargs):
        if mode == 'train':
            return self.train_iter(*inputs, **kwargs)
        elif mode == 'extract':
            return self.backbone(*inputs)
        else:
            raise Exception("No such mode: {}".format(mode))
