"""Trainer surface of the v1.1.0 tree (passl_v110/engine/trainer.py:72-337 + hooks/optimizer_hook.py:20-48 +
hooks/log_hook.py): `Trainer(cfg).train()` iterates `outputs = model(*data)`, `outputs['loss'].backward()`, gradient sync,
optimizer step, LR schedule and the `ips: ... images/sec` log line the reference's CI greps (passl/engine/loops/loop.py:102-118).

Only what the hot path needs is here: synthetic or user-supplied iterables of (view_q, view_k) batches, the name-based model
registry, the fused optimizers, the data-parallel gradient exchange and checkpoint / resume in the spirit of
hooks/checkpoint_hook.py:22-49 (`{output_dir}/iter_{N}.pd` holding model, optimizer, lr-scheduler state and the iteration; torch
serialisation; a path ending in `.pdparams` is written / read in the reference's own container, names and layouts through
utils/checkpoint.py).  Evaluation hooks are out of scope."""
import math
import os
import sys
import time

import torch

from ..core.param_store import ParamStore
from ..distributed import get_rank, get_world_size, grad_sync, model_sync, param_sync
from ..modeling import build_model
from ..optimizer import build_lr_scheduler, build_lr_scheduler_simclr, build_optimizer


class IterLoader:
    """passl_v110/engine/trainer.py:48-69"""

    def __init__(self, dataloader, epoch=0):
        self._dataloader = dataloader
        self.iter_loader = iter(self._dataloader)
        self._epoch = epoch

    @property
    def epoch(self):
        return self._epoch

    def __next__(self):
        try:
            data = next(self.iter_loader)
        except StopIteration:
            self._epoch += 1
            self.iter_loader = iter(self._dataloader)
            data = next(self.iter_loader)
        return data

    def __len__(self):
        return len(self._dataloader)


class SyntheticTwoViews:
    """Synthetic 3x224x224 two-view batches resident on the device (the metric's data source)."""

    def __init__(self, batch_size, iters, device, size=224, seed=1234):
        self.batch_size, self.iters, self.device, self.size = batch_size, iters, device, size
        g = torch.Generator(device=device).manual_seed(seed + get_rank())
        self.a = torch.randn(batch_size, 3, size, size, device=device, generator=g)
        self.b = torch.randn(batch_size, 3, size, size, device=device, generator=g)

    def __len__(self):
        return self.iters

    def __iter__(self):
        for _ in range(self.iters):
            yield self.a, self.b


class SyntheticSingleView:
    """Synthetic single-image batches (MAE pre-training: one N(0,1) image per sample)."""

    def __init__(self, batch_size, iters, device, size=224, seed=1234):
        self.iters = iters
        g = torch.Generator(device=device).manual_seed(seed + get_rank())
        self.img = torch.randn(batch_size, 3, size, size, device=device, generator=g)

    def __len__(self):
        return self.iters

    def __iter__(self):
        for _ in range(self.iters):
            yield (self.img,)


class SyntheticImageText:
    """Synthetic (image, text) batches for the CLIP path: N(0,1) images, random token ids with the EOT id (vocab - 1) at a random
    position >= 1 (SURVEY §8d synthetic-input spec)."""

    def __init__(self, batch_size, iters, device, size=224, context_length=77, vocab_size=49408, seed=1234):
        self.iters = iters
        g = torch.Generator(device=device).manual_seed(seed + get_rank())
        self.img = torch.randn(batch_size, 3, size, size, device=device, generator=g)
        self.text = torch.randint(1, vocab_size - 1, (batch_size, context_length), device=device, generator=g)
        eot = torch.randint(1, context_length, (batch_size,), device=device, generator=g)
        self.text[torch.arange(batch_size, device=device), eot] = vocab_size - 1

    def __len__(self):
        return self.iters

    def __iter__(self):
        for _ in range(self.iters):
            yield self.img, self.text


class Trainer:
    def __init__(self, cfg, dataloader=None, device=None):
        self.cfg = cfg
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.model = build_model(dict(cfg.model)).to(self.device)
        enc = getattr(self.model, "encoder_q", None) or getattr(self.model, "encoder", None) or self.model
        store_k = None
        if hasattr(self.model, "build_param_stores"):
            self.store, store_k = self.model.build_param_stores()
        else:
            self.store = ParamStore(enc)
        model_sync(self.model, (self.store, store_k))       # parameters (stop_gradient ones too) and buffers from rank 0
        self.batch_size = cfg.dataloader.train.sampler.batch_size
        if dataloader is None:
            iters = cfg.get("total_iters", 10)
            ds = ((cfg.dataloader.train.get("dataset", {}) or {}).get("name", "")) if "dataloader" in cfg else ""
            if ds == "TextImageDataset":
                arch = cfg.model.architecture
                dataloader = SyntheticImageText(self.batch_size, iters, self.device, size=arch.image_resolution,
                                                context_length=arch.context_length, vocab_size=arch.vocab_size)
            elif (cfg.dataloader.train.get("dataset", {}) or {}).get("single_view", False):
                dataloader = SyntheticSingleView(self.batch_size, iters, self.device)
            elif cfg.dataloader.train.get("device_input_stage", False):
                # decoded uint8 images -> two augmented views on the GPU, recipe taken from the YAML's transform lists (f-2)
                from ..data import DeviceAugmentedTwoViews, SyntheticDecodedImages, build_input_stage
                dataloader = DeviceAugmentedTwoViews(SyntheticDecodedImages(self.batch_size, iters, self.device),
                                                     build_input_stage(cfg.dataloader.train.dataset))
            else:
                dataloader = SyntheticTwoViews(self.batch_size, iters, self.device)
        # LR schedule (engine/trainer.py:140-166 of the reference): epoch-denominated YAML keys become iterations through
        # iters_per_epoch = len(dataloader); the SimCLR recipe derives rate, warm-up and horizon from batch size and image count
        # engine/trainer.py:228-233 of the reference: epochs set -> total_iters = epochs * iters_per_epoch, else cfg.total_iters;
        # an IterLoader wraps around the loader, so the loader's length is one epoch
        self._epochs_cfg = cfg.get("epochs", None)
        self.iters_per_epoch = cfg.get("iters_per_epoch", None) or len(dataloader)
        self.epochs = self._epochs_cfg or 1
        opt_cfg = dict(cfg.optimizer)
        lr_cfg = dict(cfg.get("lr_scheduler", {}) or {})
        self.lr_scheduler = None
        if lr_cfg:
            if cfg.get("use_simclr_iters", False):
                self.lr_scheduler = build_lr_scheduler_simclr(lr_cfg, self.iters_per_epoch, self.batch_size * 8, self.epochs, 0)
            else:
                self.lr_scheduler = build_lr_scheduler(lr_cfg, self.iters_per_epoch)
            opt_cfg["lr"] = self.lr_scheduler()      # the scheduler object is the optimizer's learning rate (solver/builder.py:88)
        self.optimizer = build_optimizer(opt_cfg, self.store)
        self.dataloader = dataloader
        self.log_interval = (cfg.get("log_config", {}) or {}).get("interval", 10)
        self.current_iter = 0
        self.outputs = None
        self.output_dir = cfg.get("output_dir", None)
        self.checkpoint_interval = (cfg.get("checkpoint", {}) or {}).get("interval", 0)

    # -- checkpoint / resume (rank 0 writes; every rank can load) ----------------------------------------------------------------
    def save(self, path=None, paddle_format=False):
        if path is None:
            path = os.path.join(self.output_dir or ".", "iter_%d.pd" % self.current_iter)
        if path.endswith(".pdparams") or paddle_format:
            # the reference's own containers / parameter names / layouts, see utils/checkpoint.py
            if get_rank() == 0:
                from ..utils import checkpoint as C
                os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
                if path.endswith(".pdparams"):
                    C.save_pdparams(C.to_paddle_state(self.model), path)
                else:                                    # v110 `epoch_N.pd` (hooks/checkpoint_hook.py:22-49), without optimizer state
                    C.save_v110_checkpoint(path, self.model, self.current_iter // max(1, self.iters_per_epoch) + 1, self.lr_scheduler)
            return path
        if get_rank() == 0:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            torch.save({"iter": self.current_iter, "state_dict": self.model.state_dict(),
                        "optimizer": self.optimizer.state_dict(),
                        "lr_scheduler": None if self.lr_scheduler is None else {"last_epoch": self.lr_scheduler.last_epoch}}, path)
        return path

    def _weights_changed(self):
        for st in (getattr(self.model, "_stores", None) or (self.store,)):
            if st is not None:
                st.refresh_bf16()                        # parameters are views of the flat fp32 master: refresh the bf16 mirror
        if hasattr(self.model, "_queue_bf16") and self.model._queue_bf16 is not None:
            from .. import kernels as K
            self.model._queue_bf16 = K.cast_bf16(self.model.queue)

    def load(self, weight_path):
        """Weights only (engine/trainer.py:433-444): a .pdparams file, a v110 checkpoint (its 'state_dict' entry) or this package's
        own checkpoint; iteration, schedule and optimizer state are left as constructed."""
        from ..utils import checkpoint as C
        if weight_path.endswith(".pdparams") or C.is_paddle_pickle(weight_path):
            C.load_paddle_state(self.model, C.load_v110_checkpoint(weight_path)["state_dict"])
        else:
            self.model.load_state_dict(torch.load(weight_path, map_location=self.device)["state_dict"])
        self._weights_changed()

    def resume(self, path):
        from ..utils import checkpoint as C
        if path.endswith(".pdparams") or C.is_paddle_pickle(path):
            # the reference's containers: a bare .pdparams weights file, or a v110 `epoch_N.pd` training checkpoint
            # (engine/trainer.py:419-437).  Weights, epoch and LR-schedule position are restored; optimizer moments are not.
            ck = C.load_v110_checkpoint(path)
            C.load_paddle_state(self.model, ck["state_dict"])
            self._weights_changed()
            if ck.get("epoch") is not None:
                self.current_iter = (ck["epoch"] - 1) * self.iters_per_epoch
            if self.lr_scheduler is not None and ck.get("lr_scheduler"):
                self.lr_scheduler.set_state_dict(ck["lr_scheduler"])
                self.optimizer.set_lr(self.lr_scheduler())
            return
        ck = torch.load(path, map_location=self.device)
        self.model.load_state_dict(ck["state_dict"])
        self._weights_changed()
        self.optimizer.set_state_dict(ck["optimizer"])
        if self.lr_scheduler is not None and ck.get("lr_scheduler"):
            self.lr_scheduler.set_state_dict(ck["lr_scheduler"])
        self.current_iter = int(ck["iter"])

    @property
    def total_iters(self):
        """engine/trainer.py:228-233 of the reference: `epochs` set -> epochs * iters_per_epoch (one pass of the loader = one epoch,
        the IterLoader wraps around), else cfg.total_iters"""
        ipe = self.cfg.get("iters_per_epoch", None) or len(self.dataloader)
        return self._epochs_cfg * ipe if self._epochs_cfg else int(self.cfg.get("total_iters", len(self.dataloader)))

    def train(self):
        from ..utils.profiler import StepProfiler
        profiler = StepProfiler(self.cfg.get("profiler_options", None))       # -p "batch_range=[a, b]; ..." (tools/train.py:30)
        loader = IterLoader(self.dataloader)
        total = self.total_iters
        self.iters_per_epoch = self.cfg.get("iters_per_epoch", None) or len(self.dataloader)
        t0, seen = time.time(), 0
        while self.current_iter < total:
            profiler.step()
            data = next(loader)
            self.optimizer.clear_grad()
            self.outputs = self.model(*data, total_iters=total, current_iter=self.current_iter)
            self.outputs['loss'].backward()                       # OptimizerHook.train_iter_end (optimizer_hook.py:25-48)
            grad_sync(self.store)
            self.optimizer.step()
            if self.lr_scheduler is not None:                      # LRSchedulerHook (by iteration)
                self.optimizer.set_lr(self.lr_scheduler.step())
            self.current_iter += 1
            seen += self.batch_size * get_world_size()
            if self.current_iter % self.log_interval == 0 and get_rank() == 0:
                torch.cuda.synchronize()
                dt = time.time() - t0
                loss_value = float(self.outputs['loss'].detach())
                msg = "[Train][Iter: {}/{}] lr: {:.5f}, loss: {:.5f}, batch_cost: {:.5f}s, ips: {:.5f} images/sec".format(
                    self.current_iter, total, self.optimizer.get_lr(), loss_value, dt / self.log_interval, seen / dt)
                print(msg, flush=True)
                if not math.isfinite(loss_value):            # tasks/ssl/mae/engine_pretrain.py:73-75; checked where the loss is
                    print("Loss is {}, stopping training".format(loss_value), flush=True)      # read back anyway (log steps)
                    sys.exit(1)
                t0, seen = time.time(), 0
            if self.checkpoint_interval and self.current_iter % self.checkpoint_interval == 0:
                self.save()
        return self.outputs
