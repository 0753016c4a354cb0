"""A minimal torch-CPU-backed stand-in for the `paddle` package — TEST INFRASTRUCTURE ONLY.

Purpose: execute the reference's OWN Python source for the loss heads / queue / MAE loss (read from /root/reference at
golden-generation time, never copied) so that oracle/ can be pinned against outputs of the reference code itself.
Only the Paddle ops those few functions touch are provided; each restates the documented Paddle semantics
(paddle 2.4 API docs) on float64 torch tensors.  PaddlePaddle itself is not installable in this environment.
"""
import sys
import types

import torch


class _Permissive(types.ModuleType):
    """Module whose unknown attributes resolve to inert placeholders so unrelated imports in reference files succeed."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Permissive(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return self

    def __mro_entries__(self, bases):
        return (object,)


def _axis_kw(kw):
    if "axis" in kw:
        kw["dim"] = kw.pop("axis")
    if "keepdim" not in kw and "keepdims" in kw:
        kw["keepdim"] = kw.pop("keepdims")
    return kw


_orig_transpose = torch.Tensor.transpose


def _transpose(self, *args, **kw):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        return self.permute(*args[0])
    return _orig_transpose(self, *args, **kw)


def install():
    """Create and register the shim as sys.modules['paddle'] (+ submodules). Returns the module."""
    torch.Tensor.transpose = _transpose           # paddle: x.transpose([1, 0])
    torch.Tensor.astype = lambda self, dt: self.to(_dtype(dt))
    torch.Tensor.numpy_ = torch.Tensor.numpy
    _tile = torch.Tensor.tile
    torch.Tensor.tile = lambda self, *reps: _tile(self, *(reps[0] if len(reps) == 1 and isinstance(reps[0], (tuple, list)) else reps))

    paddle = _Permissive("paddle")
    nn = _Permissive("paddle.nn")
    F = _Permissive("paddle.nn.functional")
    fluid = _Permissive("paddle.fluid")
    layers = _Permissive("paddle.fluid.layers")
    dist = _Permissive("paddle.distributed")

    def _dtype(d):
        if isinstance(d, torch.dtype):
            return d
        return {"float32": torch.float64, "float64": torch.float64, "int64": torch.int64, "int32": torch.int32,
                "bool": torch.bool}[str(d).replace("paddle.", "")]

    globals()["_dtype"] = _dtype

    # ---- tensor creation / manipulation (float32 is promoted to float64: the oracle is an fp64 reference) ----
    paddle.float32, paddle.int64, paddle.int32 = torch.float64, torch.int64, torch.int32
    paddle.to_tensor = lambda x, dtype=None, **k: torch.as_tensor(x, dtype=_dtype(dtype) if dtype else None)
    paddle.zeros = lambda shape, dtype="float32": torch.zeros(tuple(shape), dtype=_dtype(dtype))
    paddle.ones = lambda shape, dtype="float32": torch.ones(tuple(shape), dtype=_dtype(dtype))
    paddle.concat = lambda xs, axis=0: torch.cat(list(xs), dim=axis)
    paddle.reshape = lambda x, shape: x.reshape(tuple(shape))
    paddle.unsqueeze = lambda x, axis: x.unsqueeze(axis)
    paddle.cast = lambda x, dtype: x.to(_dtype(dtype))
    paddle.sum = lambda x, axis=None, keepdim=False: x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdim)
    paddle.argmax = lambda x, axis=None: x.argmax(dim=axis)
    paddle.argsort = lambda x, axis=-1: torch.argsort(x, dim=axis, stable=True)
    paddle.einsum = torch.einsum

    def arange(start, end=None, step=1, dtype="int64"):
        if end is None:
            start, end = 0, start
        return torch.arange(start, end, step, dtype=_dtype(dtype))
    paddle.arange = arange

    def matmul(x, y, transpose_x=False, transpose_y=False):
        if transpose_x:
            x = x.transpose(-1, -2)
        if transpose_y:
            y = y.transpose(-1, -2)
        return x @ y
    paddle.matmul = matmul

    class no_grad:
        def __call__(self, fn=None):
            return torch.no_grad()(fn) if fn is not None else self

        def __enter__(self):
            self._g = torch.no_grad()
            return self._g.__enter__()

        def __exit__(self, *a):
            return self._g.__exit__(*a)
    paddle.no_grad = no_grad()

    # ---- nn ----
    class Layer(torch.nn.Module):
        def register_buffer(self, name, tensor, persistable=True):
            super().register_buffer(name, tensor)
    nn.Layer = Layer
    nn.Sequential = torch.nn.Sequential

    class CrossEntropyLoss(Layer):
        """paddle.nn.CrossEntropyLoss(): softmax + NLL on hard labels, reduction='mean'."""

        def forward(self, logits, labels):
            lsm = torch.log_softmax(logits.double(), dim=-1)
            return -lsm.gather(1, labels.reshape(-1, 1).long()).mean()
    nn.CrossEntropyLoss = CrossEntropyLoss

    F.one_hot = lambda x, num_classes: torch.nn.functional.one_hot(x.long(), num_classes).double()
    F.softmax = lambda x, axis=-1: torch.softmax(x, dim=axis)
    F.log_softmax = lambda x, axis=-1: torch.log_softmax(x, dim=axis)
    F.normalize = lambda x, p=2, axis=1, epsilon=1e-12: x / x.norm(p=p, dim=axis, keepdim=True).clamp_min(epsilon)

    def softmax_with_cross_entropy(logits, label, soft_label=False, axis=-1):
        lsm = torch.log_softmax(logits, dim=axis)
        if soft_label:
            return -(label * lsm).sum(dim=axis, keepdim=True)
        return -lsm.gather(axis, label.long())
    F.softmax_with_cross_entropy = softmax_with_cross_entropy

    def kl_div(input, label, reduction="mean"):
        t = torch.where(label > 0, label * (torch.log(label.clamp_min(1e-300)) - input), torch.zeros_like(label))
        if reduction == "batchmean":
            return t.sum() / input.shape[0]
        if reduction == "sum":
            return t.sum()
        if reduction == "mean":
            return t.mean()
        return t
    F.kl_div = kl_div

    layers.reduce_mean = lambda x: x.mean()

    def accuracy(input, label, k=1):
        topk = input.topk(k, dim=1).indices
        return (topk == label.reshape(-1, 1)).any(dim=1).double().mean()
    layers.accuracy = accuracy
    layers.l2_normalize = lambda x, axis, epsilon=1e-12: x / torch.sqrt((x * x).sum(dim=axis, keepdim=True) + epsilon)

    dist.get_world_size = lambda: 1
    dist.get_rank = lambda: 0

    nn.functional = F
    paddle.nn = nn
    fluid.layers = layers
    paddle.fluid = fluid
    paddle.distributed = dist
    for name, mod in [("paddle", paddle), ("paddle.nn", nn), ("paddle.nn.functional", F), ("paddle.fluid", fluid),
                      ("paddle.fluid.layers", layers), ("paddle.distributed", dist)]:
        sys.modules[name] = mod
    return paddle


class _PaddleFinder:
    """Any `paddle.<anything>` import that the shim does not define resolves to an inert placeholder module."""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        if not fullname.startswith("paddle."):
            return None
        import importlib.machinery

        class _Loader:
            @staticmethod
            def create_module(spec):
                m = _Permissive(spec.name)
                m.__path__ = []
                return m

            @staticmethod
            def exec_module(module):
                pass
        return importlib.machinery.ModuleSpec(fullname, _Loader(), is_package=True)


def install_finder():
    sys.modules["paddle"].__path__ = []
    if not any(isinstance(f, type) and f is _PaddleFinder for f in sys.meta_path):
        sys.meta_path.append(_PaddleFinder)


def fake_package(name, path):
    """Register an empty package whose __path__ points at a reference directory (its __init__ is NOT executed)."""
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m
