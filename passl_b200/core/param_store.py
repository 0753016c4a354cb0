"""Flat parameter storage: one fp32 master buffer, one bf16 compute mirror, one fp32 gradient buffer per model.

B200-native take on the reference's tensor fusion (passl/core/param_fuse.py:418-505, passl/optimizer/__init__.py:172-182):
every parameter becomes a view into a flat buffer, so the optimizer step, the momentum-encoder EMA (moco.py:82-90), the
bf16 refresh and the data-parallel gradient all-reduce (passl/core/sync_utils.py:18-43) are each ONE kernel / collective
instead of a Python loop over ~160 tensors.  Tensors start at multiples of 1024 elements (TMA needs 16-byte aligned bases;
the LARS kernel needs blocks that never straddle two tensors).
"""
import torch

ALIGN = 1024


class ParamStore:
    def __init__(self, module, device=None, with_grad=True):
        params = [p for p in module.parameters()]
        assert params, "module has no parameters"
        device = device or params[0].device
        self.params = params
        self.names = [n for n, _ in module.named_parameters()]
        self.module = module                         # optimizers derive Paddle-style auto names from the module tree
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.master = torch.zeros(off, dtype=torch.float32, device=device)
        self.bf16 = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.grad = torch.zeros(off, dtype=torch.float32, device=device) if with_grad else None
        seg = torch.empty(off // ALIGN, dtype=torch.int32)
        for i, (p, o) in enumerate(zip(params, self.offsets)):
            n = p.numel()
            self.master[o:o + n].copy_(p.detach().reshape(-1).float())
            p.data = self.master[o:o + n].view(p.shape)
            p.bf16 = self.bf16[o:o + n].view(p.shape)
            if with_grad and p.requires_grad:
                p.grad = self.grad[o:o + n].view(p.shape)
            seg[o // ALIGN:(o + (n + ALIGN - 1) // ALIGN * ALIGN) // ALIGN] = i
        self.block_seg = seg.to(device)
        self.refresh_bf16()
        module._param_store = self

    def refresh_bf16(self):
        from .. import kernels as K
        if self.master.is_cuda:
            K.cast_bf16(self.master, out=self.bf16)
        else:  # CPU construction path (unit tests of the host logic only)
            self.bf16.copy_(self.master)

    def zero_grad(self):
        if self.grad is not None:
            self.grad.zero_()

    def segment_values(self, fn, dtype=torch.float32):
        """Per-tensor scalar table (e.g. weight-decay mask) on the store's device: fn(name, param) -> float."""
        return torch.tensor([fn(n, p) for n, p in zip(self.names, self.params)], dtype=dtype, device=self.master.device)

    def copy_from(self, other):
        """Parameter-wise copy (MoCo key-encoder initialisation: param_k.set_value(param_q), moco.py:65-67)."""
        assert self.numel == other.numel
        self.master.copy_(other.master)
        self.bf16.copy_(other.bf16)


def compute_copy(p):
    """bf16 view used by the tcgen05 kernels (falls back to an on-the-fly cast when no ParamStore was built)."""
    b = getattr(p, "bf16", None)
    if b is not None:
        return b
    from .. import kernels as K
    return K.cast_bf16(p.detach().contiguous())


def grad_buffer(p):
    """fp32 gradient accumulator of a parameter (allocated on demand when no ParamStore was built)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32)
    return p.grad
