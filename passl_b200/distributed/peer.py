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
