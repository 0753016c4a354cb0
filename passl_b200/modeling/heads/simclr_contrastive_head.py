"""SimCLRContrastiveHead (passl_v110/modeling/heads/simclr_contrastive_head.py:26-102): NT-Xent on [ab|aa], [ba|bb] with
the aa/bb diagonals masked, plus 3 x CO2 (KL(a||b) + KL(b||a)); returns {'loss', 'acc1'}.
`multi_rank=True` turns on the all-gathered negatives of BASELINE config 2 (the reference flag is declared but unused)."""
import torch.nn as nn

from ...distributed import all_gather, get_rank, get_world_size
from ...loss.simclr import ntxent_co2
from ..registry import HEADS


@HEADS.register()
class SimCLRContrastiveHead(nn.Module):
    def __init__(self, temperature=0.5, return_accuracy=True, multi_rank=False, co2_weight=3.0):
        super().__init__()
        self.temperature = temperature
        self.return_accuracy = return_accuracy
        self.multi_rank = multi_rank
        self.co2_weight = co2_weight

    def forward_fused(self, con, n):
        """con = [hidden1; hidden2] fp32 [2n, d] straight from the encoder (no split / re-concat round trip)."""
        gather = all_gather if (self.multi_rank and get_world_size() > 1) else None
        loss, acc1 = ntxent_co2(con, n, self.temperature, self.co2_weight, gather=gather, rank=get_rank())
        return dict(loss=loss, acc1=acc1)

    def forward(self, pos, neg):
        """Reference signature: (hidden1 [n, d], hidden2 [n, d])."""
        import torch
        return self.forward_fused(torch.cat([pos, neg], dim=0), pos.shape[0])
