"""Embedding all-gather / reduce-scatter over NVLink peer memory (csrc/peer.cu) for ranks on one node.

Each rank allocates one device buffer (`passl_b200_peer_buffer_create`: plain cudaMalloc, the library's only own allocation) holding
two slots of {gather shard, gathered-gradient buffer, flag rows}; its 64-byte CUDA IPC handle is exchanged once with
`all_gather_object` and every rank maps every peer's buffer with its own device as accessor (`passl_b200_peer_buffer_open`).  An exchange is then ONE kernel launch on the compute stream: signal through flags in peer memory, read the peers' shards
with P2P loads.  No NCCL call, no host synchronisation.  `PeerExchange.all_gather` is differentiable (backward = the peer
reduce-scatter), mirroring passl/distributed/nn/functional.py:100-127.

Falls back to nothing: constructing it outside an initialised NCCL process group with world_size > 1 raises.
"""
import ctypes

import torch
import torch.distributed as dist

from .. import _lib
from ..kernels import _ptr, _stream


def _align(n, a=256):
    return (n + a - 1) // a * a


class PeerExchange:
    def __init__(self, shard_rows, dim, device=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
            raise RuntimeError("PeerExchange needs an initialised process group with world_size > 1")
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.n, self.d = int(shard_rows), int(dim)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        assert (self.n * self.d * 4) % 16 == 0
        lib = _lib.load()
        shard_b = self.n * self.d * 4
        flag_b = _align(self.world * 4)
        # slot layout: [gather shard | gradient of the gathered tensor (world shards) | gather flags | reduce-scatter flags]
        self.off_g, self.off_r = 0, _align(shard_b)
        self.off_fg = self.off_r + _align(shard_b * self.world)
        self.off_fr = self.off_fg + flag_b
        self.slot_b = self.off_fr + flag_b
        with torch.cuda.device(self.device):
            base = ctypes.c_void_p()
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(lib.passl_b200_peer_buffer_create(2 * self.slot_b, ctypes.byref(base), handle), "peer_buffer_create")
            self._base = base.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle))
            self._mapped, bases = [], []
            for q, h in enumerate(handles):
                if q == self.rank:
                    bases.append(self._base)
                    continue
                m = ctypes.c_void_p()
                hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(lib.passl_b200_peer_buffer_open(hb, ctypes.byref(m)), "peer_buffer_open")
                self._mapped.append(m.value)
                bases.append(m.value)
        dist.barrier()
        arr = ctypes.c_void_p * self.world

        def ptrs(off):
            return [arr(*[b + s * self.slot_b + off for b in bases]) for s in range(2)]
        self._g_data, self._g_flags = ptrs(self.off_g), ptrs(self.off_fg)
        self._r_data, self._r_flags = ptrs(self.off_r), ptrs(self.off_fr)
        self._g_step = self._r_step = 0

    def close(self):
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier()
        with torch.cuda.device(self.device):
            for m in self._mapped:
                lib.passl_b200_peer_buffer_close(ctypes.c_void_p(m))
            self._mapped = []
            dist.barrier()
            if self._base:
                lib.passl_b200_peer_buffer_destroy(ctypes.c_void_p(self._base))
                self._base = None

    @torch.no_grad()
    def gather(self, x):
        """x fp32 [n, d] (this rank's shard) -> [world*n, d], rank-major (concat_all_gather semantics)."""
        assert x.shape == (self.n, self.d) and x.dtype == torch.float32 and x.is_contiguous()
        lib = _lib.load()
        slot, epoch = self._g_step & 1, (self._g_step >> 1) + 1
        self._g_step += 1
        out = torch.empty((self.world * self.n, self.d), dtype=torch.float32, device=self.device)
        _lib.check(lib.passl_b200_peer_allgather(_ptr(x), self._g_data[slot], self._g_flags[slot], _ptr(out), self.n * self.d * 4,
                                                 self.rank, self.world, epoch, _stream()), "peer_allgather")
        return out

    @torch.no_grad()
    def reduce_scatter(self, grad_all):
        """grad_all fp32 [world*n, d] (this rank's gradient of the gathered tensor) -> sum over ranks of their rows [rank*n, (rank+1)*n)."""
        assert grad_all.shape == (self.world * self.n, self.d) and grad_all.dtype == torch.float32 and grad_all.is_contiguous()
        lib = _lib.load()
        slot, epoch = self._r_step & 1, (self._r_step >> 1) + 1
        self._r_step += 1
        out = torch.empty((self.n, self.d), dtype=torch.float32, device=self.device)
        _lib.check(lib.passl_b200_peer_reduce_scatter_f32(_ptr(grad_all), self._r_data[slot], self._r_flags[slot], _ptr(out),
                                                          self.n * self.d, self.rank, self.world, epoch, _stream()),
                   "peer_reduce_scatter_f32")
        return out

    def all_gather(self, x):
        """Differentiable gather along dim 0 (same contract as passl_b200.distributed.all_gather)."""
        return _PeerAllGather.apply(x, self)


class _PeerAllGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ex):
        ctx.ex = ex
        return ex.gather(x.contiguous().float())

    @staticmethod
    def backward(ctx, grad):
        return ctx.ex.reduce_scatter(grad.contiguous().float()), None


class PeerKeyShards:
    """Key shards for the fused gathered-key InfoNCE (csrc/infonce_tc.cu peer mode): every rank keeps its bf16 keys [n, d] in a
    peer-visible buffer; `publish(k)` writes them and raises the flags; the loss kernels of all ranks then read every shard in place
    through one TMA tensor map per rank — the all-gather of mocov3.py:173-185 never happens.  Two slots alternate (same protocol
    and safety argument as PeerExchange: a rank is at most one epoch ahead of the slowest peer)."""

    def __init__(self, shard_rows, dim, device=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
            raise RuntimeError("PeerKeyShards needs an initialised process group with world_size > 1")
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        assert self.world <= 8 and shard_rows % 64 == 0 and dim % 64 == 0
        self.n, self.d = int(shard_rows), int(dim)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        shard_b = _align(self.n * self.d * 2)
        flag_b = _align(self.world * 4)
        self.off_f = shard_b
        self.slot_b = shard_b + flag_b
        with torch.cuda.device(self.device):
            base = ctypes.c_void_p()
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(lib.passl_b200_peer_buffer_create(2 * self.slot_b, ctypes.byref(base), handle), "peer_buffer_create")
            self._base = base.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle))
            self._mapped, bases = [], []
            for q, h in enumerate(handles):
                if q == self.rank:
                    bases.append(self._base)
                    continue
                m = ctypes.c_void_p()
                hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(lib.passl_b200_peer_buffer_open(hb, ctypes.byref(m)), "peer_buffer_open")
                self._mapped.append(m.value)
                bases.append(m.value)
        dist.barrier()
        arr = ctypes.c_void_p * self.world
        self._data = [arr(*[b + s * self.slot_b for b in bases]) for s in range(2)]
        self._flags = [arr(*[b + s * self.slot_b + self.off_f for b in bases]) for s in range(2)]
        self._my_flags = [self._base + s * self.slot_b + self.off_f for s in range(2)]
        self._done = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._step = 0
        self._cur = None

    def close(self):
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier()
        with torch.cuda.device(self.device):
            for m in self._mapped:
                lib.passl_b200_peer_buffer_close(ctypes.c_void_p(m))
            self._mapped = []
            dist.barrier()
            if self._base:
                lib.passl_b200_peer_buffer_destroy(ctypes.c_void_p(self._base))
                self._base = None

    @torch.no_grad()
    def publish(self, k):
        """k fp32 [n, d]: this rank's (normalised) keys of the step.  Returns the (slot, epoch) token the loss kernels use."""
        assert k.shape == (self.n, self.d) and k.dtype == torch.float32 and k.is_contiguous()
        lib = _lib.load()
        slot, epoch = self._step & 1, (self._step >> 1) + 1
        self._step += 1
        _lib.check(lib.passl_b200_peer_publish_keys_bf16(_ptr(k), self.n, self.d, self._data[slot], self._flags[slot], self.rank,
                                                         self.world, epoch, _ptr(self._done), _stream()), "peer_publish_keys_bf16")
        self._cur = (slot, epoch)
        return self._cur

    def infonce_fwd(self, q_bf16, label, scale, loss_scale=1.0, token=None):
        from ..kernels import _infonce_state
        lib = _lib.load()
        slot, epoch = token or self._cur
        N, D = q_bf16.shape
        assert D == self.d and q_bf16.dtype == torch.bfloat16 and q_bf16.is_contiguous()
        dev = q_bf16.device
        lse = torch.empty(N, dtype=torch.float32, device=dev)
        tgt = torch.empty(N, dtype=torch.float32, device=dev)
        out = torch.empty(3, dtype=torch.float32, device=dev)
        ws = _infonce_state(lib.passl_b200_infonce_tc_workspace_bytes(N, self.world * self.n, D), dev, N)
        _lib.check(lib.passl_b200_infonce_tc_fwd_peer(_ptr(q_bf16), self._data[slot], ctypes.c_void_p(self._my_flags[slot]), self.world,
                                                      self.n, epoch, _ptr(label), None, float(scale), float(loss_scale), N, D, _ptr(lse),
                                                      _ptr(tgt), None, _ptr(out), _ptr(ws), ws.numel(), _stream()), "infonce_tc_fwd_peer")
        return out, lse, tgt

    def infonce_bwd(self, q_bf16, label, lse, tgt, scale, loss_scale=1.0, dloss=None, token=None):
        lib = _lib.load()
        slot, epoch = token or self._cur
        N, D = q_bf16.shape
        dq = torch.empty((N, D), dtype=torch.float32, device=q_bf16.device)
        _lib.check(lib.passl_b200_infonce_tc_bwd_peer(_ptr(q_bf16), self._data[slot], ctypes.c_void_p(self._my_flags[slot]), self.world,
                                                      self.n, epoch, _ptr(label), None, float(scale), float(loss_scale), N, D, _ptr(lse),
                                                      _ptr(tgt), _ptr(dloss), _ptr(dq), _stream()), "infonce_tc_bwd_peer")
        return dq


class _PeerInfoNCE(torch.autograd.Function):
    """loss_scale * mean CE of q . [k_rank0; k_rank1; ...]^T * scale with labels arange(N) + N*rank (mocov3.py:187-198): forward
    and backward read the peers' key shards in place; only q receives a gradient (the keys are no-grad in the reference)."""

    @staticmethod
    def forward(ctx, q, k_local, shards, scale, loss_scale):
        from .. import kernels as K
        token = shards.publish(k_local.detach().contiguous().float())
        qb = K.cast_bf16(q.contiguous())
        N = q.shape[0]
        label = torch.arange(N, device=q.device, dtype=torch.int64) + N * shards.rank
        out, lse, tgt = shards.infonce_fwd(qb, label, scale, loss_scale, token)
        ctx.saved = (qb, label, lse, tgt, token)
        ctx.shards, ctx.scale, ctx.loss_scale = shards, scale, loss_scale
        ctx.mark_non_differentiable(out[1], out[2])
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, dloss, _a1, _a5):
        qb, label, lse, tgt, token = ctx.saved
        dq = ctx.shards.infonce_bwd(qb, label, lse, tgt, ctx.scale, ctx.loss_scale, dloss.contiguous().float().reshape(1), token)
        return dq, None, None, None, None


def peer_gathered_infonce(q, k_local, shards, scale, loss_scale=1.0):
    """(loss, acc1, acc5) of the gathered-key InfoNCE without materialising the gathered keys (PeerKeyShards)."""
    return _PeerInfoNCE.apply(q, k_local, shards, scale, loss_scale)
