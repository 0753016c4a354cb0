"""MoCo v1/v2 (passl_v110/modeling/architectures/moco.py:25-210) on the B200 kernels.

Same constructor and forward contract: ``MoCo(backbone, neck, head, dim=128, K=65536, m=0.999, T=0.07)``,
``model(img_q, img_k, mode='train') -> {'loss','acc1','acc5'}``.  B200-native differences (results identical):
  * queue is an on-device ring buffer stored key-major [K, D] (reference: [D, K]); `queue_ptr` int64[1] stays on the
    device — no `int(queue_ptr[0])` host sync (moco.py:98), no 33 MB `queue.clone()` per step (moco.py:180): the enqueue
    of step t is deferred to the start of step t+1 (after backward(t) has consumed the old rows).
  * momentum update (moco.py:82-90) is ONE kernel over the flat parameter buffer instead of ~165 assigns.
  * the key encoder's BatchNorm uses global statistics (freeze_batchnorm_statictis, moco.py:69-74), so the shuffle-BN
    all-gather/permute/un-permute (moco.py:107-152) is a mathematical no-op on k and is elided at world_size 1.
"""
import torch
import torch.nn as nn

from ... import kernels as K
from ...core.param_store import ParamStore
from ...loss.contrastive import normalize
from ...distributed import concat_all_gather, get_world_size
from ..registry import MODELS, build_backbone, build_neck, build_head


def freeze_batchnorm_statictis(layer):
    """passl_v110/modules/freeze.py:17-23"""
    from ...nn.layers import BatchNormState
    for m in layer.modules():
        if isinstance(m, BatchNormState):
            m.use_global_stats = True


@MODELS.register()
class MoCo(nn.Module):
    def __init__(self, backbone, neck=None, head=None, dim=128, K=65536, m=0.999, T=0.07, queue_dtype="bf16",
                 literal_shuffle_bn=False):
        super().__init__()
        self.K, self.m, self.T = K, m, T
        self.literal_shuffle_bn = literal_shuffle_bn    # run the reference's shuffle protocol even though it cannot change k
        self.encoder_q = nn.Sequential(build_backbone(backbone), build_neck(neck))
        self.encoder_k = nn.Sequential(build_backbone(backbone), build_neck(neck))
        self.backbone = self.encoder_q[0]
        self.head = build_head(head)
        for pq, pk in zip(self.encoder_q.parameters(), self.encoder_k.parameters()):
            pk.data.copy_(pq.data)          # initialize
            pk.requires_grad = False        # not update by gradient
        freeze_batchnorm_statictis(self.encoder_k)
        queue = torch.nn.functional.normalize(torch.randn(dim, K), dim=0)   # moco.py:77-78 (reference layout [dim, K])
        self.register_buffer("queue", queue.t().contiguous())               # stored key-major [K, dim]
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.int64))
        self.queue_dtype = queue_dtype
        self._queue_bf16 = None
        self._pending_keys = None
        self._stores = None

    # -- flat parameter storage -------------------------------------------------------------------------------
    def build_param_stores(self):
        """Call once after .cuda(): flat fp32/bf16/grad buffers for encoder_q, flat fp32/bf16 for encoder_k."""
        sq = ParamStore(self.encoder_q, with_grad=True)
        sk = ParamStore(self.encoder_k, with_grad=False)
        self._stores = (sq, sk)
        if self.queue_dtype == "bf16":
            self._queue_bf16 = K.cast_bf16(self.queue)
        return sq, sk

    @torch.no_grad()
    def refresh_buffer_mirrors(self):
        """after the queue buffer was overwritten from outside (initial broadcast, checkpoint load): rebuild its bf16 mirror"""
        if self._queue_bf16 is not None:
            self._queue_bf16 = K.cast_bf16(self.queue)

    @torch.no_grad()
    def _momentum_update_key_encoder(self):
        if self._stores is None:
            self.build_param_stores()
        sq, sk = self._stores
        K.ema_update(sk.master, sq.master, self.m, k_bf16=sk.bf16)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys):
        """moco.py:92-105; keys fp32 [N, D].  The write is deferred (see module docstring)."""
        keys = concat_all_gather(keys)
        batch_size = keys.shape[0]
        assert self.K % batch_size == 0  # for simplicity
        self._pending_keys = keys.contiguous()

    @torch.no_grad()
    def flush_queue(self):
        if self._pending_keys is not None:
            K.queue_enqueue(self._pending_keys, self.queue_ptr, queue_f32=self.queue, queue_bf16=self._queue_bf16)
            self._pending_keys = None

    def state_dict(self, *a, **kw):
        self.flush_queue()
        return super().state_dict(*a, **kw)

    # -- shuffle-BN (moco.py:107-152): same protocol over torch.distributed; off by default because encoder_k normalises with
    #    global statistics, so k does not depend on which samples share a GPU -------------------------------------------------
    @torch.no_grad()
    def _batch_shuffle_ddp(self, x):
        """all-gather the batch, rank 0 draws the permutation and broadcasts it, every rank keeps its slice of the shuffle."""
        import torch.distributed as dist
        from ...distributed import get_rank
        batch_size_this = x.shape[0]
        x_gather = concat_all_gather(x)
        batch_size_all = x_gather.shape[0]
        num_gpus = batch_size_all // batch_size_this
        idx_shuffle = torch.randperm(batch_size_all, device=x.device)
        if get_world_size() > 1:
            dist.broadcast(idx_shuffle, src=0)
        idx_unshuffle = torch.argsort(idx_shuffle)
        idx_this = idx_shuffle.view(num_gpus, -1)[get_rank()]
        return x_gather.index_select(0, idx_this), idx_unshuffle

    @torch.no_grad()
    def _batch_unshuffle_ddp(self, x, idx_unshuffle):
        from ...distributed import get_rank
        batch_size_this = x.shape[0]
        x_gather = concat_all_gather(x)
        num_gpus = x_gather.shape[0] // batch_size_this
        idx_this = idx_unshuffle.view(num_gpus, -1)[get_rank()]
        return x_gather.index_select(0, idx_this)

    def train_iter(self, *inputs, **kwargs):
        img_q, img_k = inputs
        self.flush_queue()
        q = self.encoder_q(img_q)                # queries: NxC (fp32 from the neck)
        q = normalize(q)
        with torch.no_grad():
            self._momentum_update_key_encoder()
            if self.literal_shuffle_bn and get_world_size() > 1:
                im_k, idx_unshuffle = self._batch_shuffle_ddp(img_k)
                k = normalize(self.encoder_k(im_k))
                k = self._batch_unshuffle_ddp(k, idx_unshuffle)
            else:   # elided: encoder_k BN uses global statistics, k does not depend on the batch composition
                k = self.encoder_k(img_k)
                k = normalize(k)
        queue = self._queue_bf16 if self._queue_bf16 is not None else self.queue
        outputs = self.head.forward_fused(q, k, queue)
        self._dequeue_and_enqueue(k)
        return outputs

    def forward(self, *inputs, mode='train', **kwargs):
        if mode == 'train':
            return self.train_iter(*inputs, **kwargs)
        elif mode == 'extract':
            return self.backbone(*inputs)
        else:
            raise Exception("No such mode: {}".format(mode))
