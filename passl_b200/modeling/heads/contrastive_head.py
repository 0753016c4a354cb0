"""ContrastiveHead (passl_v110/modeling/heads/contrastive_head.py:21-78).

Reference contract: ``forward(pos[N,1], neg[N,K]) -> {'loss','acc1','acc5'}`` on materialised similarities.  The
B200 path adds ``forward_fused(q, k, queue)`` which takes the embeddings and never materialises the [N, K+1] logits
(SURVEY.md §8 b "Head contract"); MoCo calls the fused entry.
"""
import torch
import torch.nn as nn

from ...loss.contrastive import moco_infonce, gathered_infonce
from ..registry import HEADS


@HEADS.register()
class ContrastiveHead(nn.Module):
    def __init__(self, temperature=0.1, return_accuracy=True, precision="bf16"):
        super().__init__()
        self.temperature = temperature
        self.return_accuracy = return_accuracy
        self.precision = precision

    def forward_fused(self, q, k, queue):
        """q, k: fp32 [N, D] normalised; queue: [K, D] key-major.  labels are all-zero int64 (positive = column 0)."""
        loss, acc1, acc5 = moco_infonce(q, k, queue, self.temperature, precision=self.precision)
        outputs = dict(loss=loss)
        if self.return_accuracy:
            outputs['acc1'] = acc1
            outputs['acc5'] = acc5
        return outputs

    def forward(self, pos, neg):
        """Reference signature on materialised similarities: treated as a 1-d embedding problem per column is not
        possible, so the logits are consumed through the same fused kernel with D = K+1 one-hot 'keys' only for tiny
        debugging sizes; the training path uses forward_fused."""
        N = pos.shape[0]
        logits = torch.cat((pos, neg), dim=1).float().contiguous()          # debugging path only
        eye = torch.eye(logits.shape[1], device=logits.device)
        pad = (-logits.shape[1]) % 128
        if pad:
            logits = torch.nn.functional.pad(logits, (0, pad))
            eye = torch.nn.functional.pad(eye, (0, pad))
        labels = torch.zeros((N,), dtype=torch.int64, device=logits.device)
        loss, acc1, acc5 = gathered_infonce(logits, eye, labels, 1.0 / self.temperature, precision="fp32")
        outputs = dict(loss=loss)
        if self.return_accuracy:
            outputs['acc1'] = acc1
            outputs['acc5'] = acc5
        return outputs
