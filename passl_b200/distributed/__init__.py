"""Data-parallel helpers over torch.distributed (NCCL on GPU, gloo in CPU tests) — the reference's
passl/distributed/{env.py,nn/functional.py}, passl/core/sync_utils.py and the `concat_all_gather` helpers
(moco.py:198-210, mocov3.py:173-185)."""
import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_initialized() else 1


def get_rank():
    return dist.get_rank() if is_initialized() else 0


@torch.no_grad()
def concat_all_gather(tensor):
    """moco.py:198-210: all_gather + concat along dim 0, no gradient.  One NCCL all_gather_into_tensor (no list+concat copy)."""
    if get_world_size() < 2:
        return tensor
    tensor = tensor.contiguous()
    out = torch.empty((get_world_size() * tensor.shape[0],) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, tensor)
    return out


class _AllGatherGrad(torch.autograd.Function):
    """Differentiable all-gather: backward = reduce-scatter of the gathered gradient
    (passl/distributed/nn/functional.py:30-42,100-127)."""

    @staticmethod
    def forward(ctx, tensor):
        return concat_all_gather(tensor)

    @staticmethod
    def backward(ctx, grad):
        if get_world_size() < 2:
            return grad
        grad = grad.contiguous()
        out = torch.empty((grad.shape[0] // get_world_size(),) + tuple(grad.shape[1:]), dtype=grad.dtype, device=grad.device)
        dist.reduce_scatter_tensor(out, grad, op=dist.ReduceOp.SUM)
        return out


def all_gather(tensor):
    """Differentiable gather along dim 0 (concatenated form of dist_F.all_gather)."""
    return _AllGatherGrad.apply(tensor)


def grad_sync(store):
    """passl/core/sync_utils.py:18-43: mean all-reduce of every gradient — here one collective on the flat buffer."""
    if get_world_size() < 2 or store.grad is None:
        return
    dist.all_reduce(store.grad, op=dist.ReduceOp.SUM)
    # the 1/nranks scale (sync_utils.py:41) is folded into the optimizer kernel's grad_scale


def param_sync(store, src=0):
    """passl/core/sync_utils.py:46-69: broadcast parameters from rank 0."""
    if get_world_size() < 2:
        return
    dist.broadcast(store.master, src=src)
    store.refresh_bf16()


def model_sync(model, stores, src=0):
    """Initial broadcast of EVERYTHING the reference's param_sync / DataParallel wrap broadcasts (passl/core/sync_utils.py:46-69):
    every parameter, stop_gradient ones included (MoCo's key encoder, MoCo v3's momentum encoder — each is a flat store here),
    and every registered buffer (BatchNorm running statistics, MoCo's queue and queue_ptr)."""
    if get_world_size() < 2:
        return
    for st in stores:
        if st is not None:
            param_sync(st, src=src)
    for b in model.buffers():
        if b.is_cuda or dist.get_backend() == "gloo":
            dist.broadcast(b, src=src)
    refresh = getattr(model, "refresh_buffer_mirrors", None)
    if refresh is not None:
        refresh()
