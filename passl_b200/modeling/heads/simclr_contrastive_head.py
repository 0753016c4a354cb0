"""SimCLRContrastiveHead (passl_v110/modeling/heads/simclr_contrastive_head.py:26-102): NT-Xent on [ab|aa], [ba|bb] with
the aa/bb diagonals masked, plus 3 x CO2 (KL(a||b) + KL(b||a)); returns {'loss', 'acc1'}.
`multi_rank=True` turns on the all-gathered negatives of BASELINE config 2 (the reference flag is declared but unused)."""
import torch.nn as nn

from ...distributed import all_gather, get_rank, get_world_size
from ...loss.simclr import ntxent_co2
from ..registry import HEADS


@HEADS.register()
class SimCLRContrastiveHead(nn.Module):
    def __init__(self, temperature=0.5, return_accuracy=True, multi_rank=False, co2_weight=3.0, peer_exchange=None):
        super().__init__()
        self.temperature = temperature
        self.return_accuracy = return_accuracy
        self.multi_rank = multi_rank
        self.co2_weight = co2_weight
        # embedding exchange: NCCL all-gather / reduce-scatter (default, works across nodes) or the single-node peer-memory kernels
        # of csrc/peer.cu (peer_exchange=True or PASSL_B200_PEER_EXCHANGE=1)
        import os
        self.peer_exchange = (os.environ.get("PASSL_B200_PEER_EXCHANGE", "0") == "1") if peer_exchange is None else bool(peer_exchange)
        self._ex = None

    def _gather_fn(self, con):
        if not (self.multi_rank and get_world_size() > 1):
            return None
        if not self.peer_exchange:
            return all_gather
        if self._ex is not None and (self._ex.n, self._ex.d) != tuple(con.shape):
            # a different shard shape (e.g. a smaller last batch): the exchange buffers were sized at construction; creating a new
            # exchange mid-step needs a collective handshake that deadlocks unless every rank changes shape at the same step —
            # take the NCCL path for this call and keep the exchange for the regular shape
            return all_gather
        if self._ex is None:
            from ...distributed.peer import PeerExchange
            self._ex = PeerExchange(con.shape[0], con.shape[1], device=con.device)
        return self._ex.all_gather

    def forward_fused(self, con, n):
        """con = [hidden1; hidden2] fp32 [2n, d] straight from the encoder (no split / re-concat round trip)."""
        gather = self._gather_fn(con)
        loss, acc1 = ntxent_co2(con, n, self.temperature, self.co2_weight, gather=gather, rank=get_rank())
        return dict(loss=loss, acc1=acc1)

    def forward(self, pos, neg):
        """Reference signature: (hidden1 [n, d], hidden2 [n, d])."""
        import torch
        return self.forward_fused(torch.cat([pos, neg], dim=0), pos.shape[0])
