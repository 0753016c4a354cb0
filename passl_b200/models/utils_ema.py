"""Shadow-weight EMA with the reference's surface (passl/models/utils/ema.py:18-97): register / update / apply_shadow / restore /
state_dict.  The reference loops over parameters in Python (one clone, two scalings and an add per tensor); here the shadow is ONE
flat fp32 buffer next to the ParamStore master and `update()` is one launch of the EMA kernel over it
(shadow = decay * shadow + (1 - decay) * param — csrc/embed.cu::ema_update, the kernel behind MoCo's key encoder).
Frozen (stop_gradient) tensors are skipped by the reference; in the flat buffer they are carried along — they never change, so
their shadow equals them — which keeps the update a single contiguous pass."""
import torch

from .. import kernels as K


class EMA:
    def __init__(self, store, decay=0.9999, thres_steps=True):
        self.store = store
        self._decay = float(decay)
        self._thres_steps = bool(thres_steps)
        self._shadow = None
        self._backup = None
        self._update_step = 0

    @torch.no_grad()
    def register(self):
        self._shadow = self.store.master.detach().clone()

    @torch.no_grad()
    def update(self):
        decay = min(self._decay, (1 + self._update_step) / (10 + self._update_step)) if self._thres_steps else self._decay
        K.ema_update(self._shadow, self.store.master, decay)
        self._update_step += 1
        return decay

    @torch.no_grad()
    def apply_shadow(self):
        assert self._shadow is not None
        self._backup = self.store.master.detach().clone()
        self.store.master.copy_(self._shadow)
        self.store.refresh_bf16()

    @torch.no_grad()
    def restore(self):
        assert self._backup is not None
        self.store.master.copy_(self._backup)
        self.store.refresh_bf16()
        self._backup = None

    @torch.no_grad()
    def state_dict(self):
        return {"shadow": self._shadow, "thres_steps": self._thres_steps, "update_step": self._update_step, "decay": self._decay}

    @torch.no_grad()
    def set_state_dict(self, state_dict):
        self._shadow = state_dict["shadow"].to(self.store.master.device)
        self._thres_steps = state_dict["thres_steps"]
        self._update_step = state_dict["update_step"]
        self._decay = state_dict["decay"]
