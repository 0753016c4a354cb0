"""GradScaler with the reference's surface (passl/core/grad_scaler.py:20-87 over paddle.amp.GradScaler): ``scale(loss)``,
``step(optimizer)``, ``update()``, dynamic loss scaling (init 2**10, x2 every `incr_every_n_steps` clean steps, x0.5 after
`decr_every_n_nan_or_inf` overflowing steps, capped at `max_loss_scaling`).

The hot path here computes in bf16 with fp32 accumulation, which needs no loss scaling: the engine builds the scaler with
``enable=False`` and every method is a pass-through, exactly as the reference's scaler behaves when fp16 is off.  When enabled, the
unscale + finite check + (optional) global-norm clip is ONE read pass over the flat gradient buffer (optimizer.GradControl) whose
result stays on the device: the fused optimizer kernel multiplies the gradient by 1/scale on the fly and skips the whole step when a
gradient is non-finite.  ``update()`` reads the found_inf flag back (one 4-byte D2H, the same sync the reference's `if found_inf`
performs) to move the scale."""
import torch


class _ScaleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss, scale):
        ctx.scale = scale
        return loss * scale

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None


class GradScaler:
    def __init__(self, enable=True, init_loss_scaling=2.0 ** 10, max_loss_scaling=2.0 ** 32, incr_ratio=2.0, decr_ratio=0.5,
                 incr_every_n_steps=1000, decr_every_n_nan_or_inf=2, use_dynamic_loss_scaling=True, no_unscale_list=None):
        self._enable = bool(enable)
        self._scale = float(init_loss_scaling) if self._enable else 1.0
        self.max_loss_scaling = float(max_loss_scaling)
        self.incr_ratio, self.decr_ratio = float(incr_ratio), float(decr_ratio)
        self.incr_every_n_steps, self.decr_every_n_nan_or_inf = int(incr_every_n_steps), int(decr_every_n_nan_or_inf)
        self.dynamic = bool(use_dynamic_loss_scaling)
        self._good, self._bad = 0, 0
        self._control = None
        if no_unscale_list:
            raise NotImplementedError("no_unscale_list is not built")

    def scale(self, loss):
        return _ScaleLoss.apply(loss, self._scale) if self._enable else loss

    def step(self, optimizer):
        if not self._enable:
            optimizer.step()
            return
        from ..optimizer import GradControl
        self._scale = min(self._scale, self.max_loss_scaling)              # grad_scaler.py:41-44
        gc = optimizer.grad_control
        if gc is None:
            gc = optimizer.grad_control = GradControl(optimizer.store)
        gc.loss_scale = self._scale
        self._control = gc
        optimizer.step()                                                    # skipped on the device when a gradient is non-finite

    def update(self):
        if not (self._enable and self.dynamic) or self._control is None:
            return
        found = bool(self._control.found_inf.item() != 0.0)
        if found:
            self._good, self._bad = 0, self._bad + 1
            if self._bad >= self.decr_every_n_nan_or_inf:
                self._scale, self._bad = max(self._scale * self.decr_ratio, 1.0), 0
        else:
            self._bad, self._good = 0, self._good + 1
            if self._good >= self.incr_every_n_steps:
                self._scale, self._good = self._scale * self.incr_ratio, 0

    def state_dict(self):
        return dict(scale=self._scale, good=self._good, bad=self._bad)

    def load_state_dict(self, st):
        self._scale, self._good, self._bad = float(st["scale"]), int(st.get("good", 0)), int(st.get("bad", 0))
