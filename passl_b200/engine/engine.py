"""The v2.5 `Engine` surface (`passl-train -c <yaml>`; tools/train.py:25-32, passl/engine/engine.py:46-358, loops/loop.py:133-375,
loops/contrastive_learning_loop.py:26-88) for the hot-path models: the YAML sections `Global`, `Model`, `Optimizer`, `LRScheduler`,
`DataLoader.Train` drive model construction, the fused optimizer, the schedule and the epoch loop.

Kept from the reference loop, in its order: `global_step += 1` -> forward / backward -> `grad_sync` -> `optimizer.step()` (which reads
the scheduler's current value) -> `clear_grad()` -> `optimizer.lr_step(global_step)` when the decay unit is the step; an epoch-unit
schedule moves after each epoch; a checkpoint every `save_interval` epochs and after the last one, under
`{output_dir}/{Model.name}/epoch_N.*`; `max_train_step` ends the run early; a non-finite loss stops it.

Not rebuilt (outside the self-supervised training path): eval / export modes, validation loops, EMA-of-weights evaluation,
VisualDL, DALI, FP16 autocast (the compute path is bf16 with fp32 master weights, so the `FP16`
section is accepted and has nothing to configure), sharding / tensor parallel strategies.  Data: synthetic two-view batches
unless a loader is passed in (no ImageNet in this environment); the `DataLoader.Train.sampler.batch_size` key is honoured.
"""
import math
import os
import pickle
import random
import sys
import time

import numpy as np
import torch

from ..core.param_store import ParamStore
from ..distributed import get_rank, get_world_size, grad_sync, model_sync
from ..models import build_model
from ..optimizer import build_lr_scheduler_v2, build_optimizer


class SyntheticTwoViewLists:
    """`[x1, x2]` batches (what the contrastive loop hands the model after dropping the label, contrastive_learning_loop.py:66-67)."""

    def __init__(self, batch_size, steps, device, size=224, seed=1234):
        self.steps = steps
        g = torch.Generator(device=device).manual_seed(seed + get_rank())
        self.a = torch.randn(batch_size, 3, size, size, device=device, generator=g)
        self.b = torch.randn(batch_size, 3, size, size, device=device, generator=g)

    def __len__(self):
        return self.steps

    def __iter__(self):
        for _ in range(self.steps):
            yield [self.a, self.b]


class Engine:
    def __init__(self, config, mode="train", device=None, dataloader=None):
        if mode != "train":
            raise NotImplementedError("Engine mode %r: only training is on the hot path" % (mode,))
        self.mode, self.config = mode, config
        G = config["Global"]
        self.print_batch_step = G.get("print_batch_step", 10)
        self.save_interval = G.get("save_interval", 1)
        self.accum_steps = int(G.get("accum_steps", 1))      # gradient merge (contrastive_learning_loop.py:35-63)
        assert self.accum_steps >= 1
        self.max_train_step = G.get("max_train_step", None)
        assert self.max_train_step is None or (isinstance(self.max_train_step, int) and self.max_train_step > 0), \
            "max_train_step must be int dtype and greater than 0"
        self.epochs = int(G["epochs"])
        self.output_dir = G.get("output_dir", "./output/")
        strategy = dict(config.get("DistributedStrategy", {}) or {})
        unsupported = [k for k, v in strategy.items() if k != "data_parallel" and v]
        if unsupported:
            raise NotImplementedError("DistributedStrategy %s: only data_parallel is built" % unsupported)
        seed = G.get("seed", False)
        if seed:                                             # engine.py:77-84: every rank its own stream
            assert isinstance(seed, int), "The 'seed' must be a integer!"
            seed += get_rank()
            torch.manual_seed(seed)
            np.random.seed(seed)
            random.seed(seed)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.model_name = config["Model"]["name"]
        self.model = build_model(dict(config["Model"])).to(self.device)
        store_k = None
        if hasattr(self.model, "build_param_stores"):
            self.store, store_k = self.model.build_param_stores()
        else:
            self.store = ParamStore(self.model)
        if self.device.type == "cuda":
            model_sync(self.model, (self.store, store_k))   # every parameter (the momentum encoder too) and buffer from rank 0
        train_cfg = (config.get("DataLoader", {}) or {}).get("Train", {}) or {}
        self.batch_size = int((train_cfg.get("sampler", {}) or {}).get("batch_size", 32))
        if dataloader is None:
            dataloader = SyntheticTwoViewLists(self.batch_size, int(train_cfg.get("synthetic_steps", 10)), self.device)
        self.train_dataloader = dataloader
        steps_per_epoch = len(dataloader)
        self.max_steps = self.epochs * steps_per_epoch       # loop.py:145-147; MoCo v3's cosine momentum runs over it
        if hasattr(self.model, "max_steps"):
            self.model.max_steps = self.max_steps
        # LR schedule and optimizer (engine.py:215-233): the scheduler's decay unit decides when it moves
        lr_cfg = config.get("LRScheduler", None)
        opt_cfg = dict(config["Optimizer"])
        self.lr_decay_unit = opt_cfg.pop("lr_decay_unit", None) or "step"
        self.lr_scheduler = None
        if lr_cfg is not None:
            self.lr_decay_unit = dict(lr_cfg).get("decay_unit", "step")
            built = build_lr_scheduler_v2(dict(lr_cfg), self.epochs, steps_per_epoch)
            if isinstance(built, (int, float)):
                opt_cfg["lr"] = float(built)
            else:
                self.lr_scheduler = built
                opt_cfg["lr"] = built.get_lr()
        elif isinstance(opt_cfg.get("lr"), dict):
            lr_inner = dict(opt_cfg["lr"])
            lr_inner["decay_unit"] = self.lr_decay_unit
            self.lr_scheduler = build_lr_scheduler_v2(lr_inner, self.epochs, steps_per_epoch)
            opt_cfg["lr"] = self.lr_scheduler.get_lr()
        for key in ("layer_decay", "param_groups", "tensor_fusion"):
            v = opt_cfg.pop(key, None)
            # falsy = not requested (the reference YAMLs spell out `tensor_fusion: False`); tensor_fusion: True is what the flat
            # ParamStore always does (param_fuse.py:418-505)
            if v and not (key == "tensor_fusion" and v is True):
                raise NotImplementedError("Optimizer.%s is not built" % key)
        if not opt_cfg.get("grad_clip"):
            opt_cfg.pop("grad_clip", None)                   # otherwise: ClipGradByGlobalNorm as one pass (optimizer.GradControl)
        self.optimizer = build_optimizer(opt_cfg, self.store)
        # the scaler of the reference loop (engine.py:180-194): a pass-through here — bf16 compute needs no loss scaling; set
        # Global.FP16-style dynamic scaling explicitly with `loss_scaling: {enable: True, ...}` to exercise the scaled path
        from ..core.grad_scaler import GradScaler
        sc = dict(G.get("loss_scaling", {}) or {})
        self.scaler = GradScaler(enable=bool(sc.pop("enable", False)), **sc)
        self.global_step, self.cur_epoch_id = 0, 0

    # -- one optimizer step (contrastive_learning_loop.py:65-88) --------------------------------------------------------------
    def train_one_step(self, batch):
        if self.lr_scheduler is not None:
            self.optimizer.set_lr(self.lr_scheduler.get_lr())        # the optimizer reads the schedule when it steps
        loss = self.forward_backward(batch)
        grad_sync(self.store)
        self.scaler.step(self.optimizer)                    # = optimizer.step() unless loss scaling is enabled
        self.scaler.update()
        self.optimizer.clear_grad()
        if self.lr_scheduler is not None and self.lr_decay_unit == "step":
            self.lr_scheduler.step(self.global_step)
        return loss

    def forward_backward(self, batch):
        """contrastive_learning_loop.py:31-63: the batch is cut into `accum_steps` sub-batches; each runs forward + backward with
        its loss divided by accum_steps, gradients add up in the flat fp32 buffer, one optimizer step follows.  The 1/accum_steps
        is folded into the optimizer's gradient multiplier (same update, no extra pass); the returned loss is the mean."""
        A = self.accum_steps
        if A == 1:
            out = self.model(batch)
            loss = out["loss"] if isinstance(out, dict) else out
            self.scaler.scale(loss).backward()
            return loss
        bs = batch[0].shape[0]
        assert bs % A == 0, ("Bad accum_steps %d for batch size %d. This may be caused by two reasons: 1) the batch size setting is "
                             "unreasonable and cannot be divisible, 2) drop_last in the sampler configuration is not set to True." % (A, bs))
        step_size = bs // A
        from .. import kernels as K
        total = None
        for idx in range(A):
            sub = [b[idx * step_size:(idx + 1) * step_size].contiguous() for b in batch]
            out = self.model(sub)
            loss = out["loss"] if isinstance(out, dict) else out
            self.scaler.scale(loss).backward()
            with torch.no_grad():
                l = loss.detach().reshape(1).float().contiguous()
                if total is None:
                    total = torch.zeros(1, dtype=torch.float32, device=l.device)
                K.axpy(total, l, 1.0 / A)
        self.optimizer.grad_scale = 1.0 / (get_world_size() * A)
        return total[0]

    def train(self):
        from ..utils.profiler import StepProfiler
        profiler = StepProfiler(self.config.get("profiler_options", None))    # config.profiler_options = args.profiler_options
        steps_per_epoch = len(self.train_dataloader)
        self.model.train()
        for epoch_id in range(self.cur_epoch_id + 1, self.epochs + 1):
            self.cur_epoch_id = epoch_id
            tic, seen = time.time(), 0
            for batch_idx, batch in enumerate(self.train_dataloader):
                if self.max_train_step is not None and self.global_step >= self.max_train_step:
                    if get_rank() == 0:
                        print("global_step({}) >= max_train_step({}), training stops early.".format(self.global_step, self.max_train_step),
                              flush=True)
                    return self.global_step
                self.global_step += 1
                profiler.step()
                loss = self.train_one_step(batch)
                seen += self.batch_size * get_world_size()
                if (batch_idx + 1) % self.print_batch_step == 0 or batch_idx + 1 == steps_per_epoch:
                    value = float(loss.detach())                      # the only device -> host read of the loop
                    if get_rank() == 0:
                        dt = time.time() - tic
                        print("[Train][Epoch {}/{}][Iter: {}/{}] lr: {:.6f}, loss: {:.5f}, ips: {:.5f} images/sec".format(
                            epoch_id, self.epochs, batch_idx + 1, steps_per_epoch, self.optimizer.get_lr(), value, seen / max(dt, 1e-9)),
                            flush=True)
                    tic, seen = time.time(), 0
                    if not math.isfinite(value):                      # engine_pretrain.py:73-75
                        print("Loss is {}, stopping training".format(value), flush=True)
                        sys.exit(1)
            if self.lr_scheduler is not None and self.lr_decay_unit == "epoch":
                self.lr_scheduler.step(epoch_id)
            if epoch_id % self.save_interval == 0 or epoch_id == self.epochs:
                self.save_checkpoint()
        return self.global_step

    # -- checkpoints (loop.py:317-340, utils/io.py:115-200): {output_dir}/{model_name}/epoch_N.* ---------------------------------
    def checkpoint_prefix(self, epoch_id=None):
        return os.path.join(self.output_dir, self.model_name, "epoch_{}".format(self.cur_epoch_id if epoch_id is None else epoch_id))

    def save_checkpoint(self):
        """epoch_N.pdparams (reference names / layouts, utils/checkpoint.py), for MoCo v3 also epoch_N_base_encoder.pdparams (the
        trunk without the projector, mocov3.py:246-260), epoch_N.pdstates (epoch, global_step, timestamp) and — in place of Paddle's
        accumulator file — epoch_N.opt.pt with this package's optimizer / schedule state."""
        if get_rank() != 0:
            return None
        from ..utils import checkpoint as C
        prefix = self.checkpoint_prefix()
        os.makedirs(os.path.dirname(prefix), exist_ok=True)
        state = C.to_paddle_state(self.model)
        C.save_pdparams(state, prefix + ".pdparams")
        if type(self.model).__name__ == "MoCoV3Pretrain":
            trunk = {k[len("base_encoder."):]: v for k, v in state.items()
                     if k.startswith("base_encoder.") and not k.startswith("base_encoder.head")}
            C.save_pdparams(trunk, prefix + "_base_encoder.pdparams")
        with open(prefix + ".pdstates", "wb") as f:
            pickle.dump({"epoch": self.cur_epoch_id, "global_step": self.global_step,
                         "timestamp": time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(time.time()))}, f, protocol=2)
        torch.save({"optimizer": self.optimizer.state_dict(),
                    "lr_scheduler": None if self.lr_scheduler is None else self.lr_scheduler.state_dict()}, prefix + ".opt.pt")
        return prefix

    def resume(self, prefix):
        """Continue from `save_checkpoint` files: weights, epoch / global_step, optimizer moments and schedule position."""
        from ..utils import checkpoint as C
        C.load_paddle_state(self.model, C.load_pdparams(prefix + ".pdparams"))
        for st in (getattr(self.model, "_stores", None) or (self.store,)):
            if st is not None and st.master.is_cuda:
                st.refresh_bf16()
        with open(prefix + ".pdstates", "rb") as f:
            meta = pickle.load(f)
        self.cur_epoch_id, self.global_step = int(meta["epoch"]), int(meta["global_step"])
        if os.path.exists(prefix + ".opt.pt"):
            ck = torch.load(prefix + ".opt.pt", map_location=self.device)
            self.optimizer.set_state_dict(ck["optimizer"])
            if self.lr_scheduler is not None and ck.get("lr_scheduler"):
                self.lr_scheduler.set_state_dict(ck["lr_scheduler"])
