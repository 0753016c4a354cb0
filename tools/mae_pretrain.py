#!/usr/bin/env python
"""MAE pre-training entry with the argument surface of the reference's tasks/ssl/mae/main_pretrain.py (hot-path subset): model,
input size, mask ratio, norm_pix_loss, weight decay, lr / blr / min_lr, warm-up epochs, epochs, per-GPU batch size, max_train_step.
Per iteration (engine_pretrain.py:52-82): half-cycle-cosine rate for the fractional epoch, forward (random masking + masked-patch
MSE), backward, AdamW with the add_weight_decay groups (1-d tensors and biases undecayed, frozen position tables skipped), and a
stop on a non-finite loss.  Data: synthetic images (no ImageNet here); one process per GPU under torch.distributed.run.

    python tools/mae_pretrain.py --model mae_vit_base_patch16 --batch_size 64 --epochs 2 --steps_per_epoch 5 --norm_pix_loss
"""
import argparse
import math
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def get_args_parser():
    p = argparse.ArgumentParser("MAE pre-training", add_help=True)
    p.add_argument("--batch_size", default=64, type=int, help="batch size per GPU")
    p.add_argument("--epochs", default=400, type=int)
    p.add_argument("--accum_iter", default=1, type=int)
    p.add_argument("--model", default="mae_vit_base_patch16", type=str)
    p.add_argument("--input_size", default=224, type=int)
    p.add_argument("--mask_ratio", default=0.75, type=float)
    p.add_argument("--norm_pix_loss", action="store_true")
    p.add_argument("--weight_decay", type=float, default=0.05)
    p.add_argument("--lr", type=float, default=None, help="absolute learning rate")
    p.add_argument("--blr", type=float, default=1e-3, help="base rate: lr = blr * total batch size / 256")
    p.add_argument("--min_lr", type=float, default=0.0)
    p.add_argument("--warmup_epochs", type=int, default=40)
    p.add_argument("--seed", default=0, type=int)
    p.add_argument("--print_freq", default=20, type=int)
    p.add_argument("--max_train_step", default=None, type=int)
    p.add_argument("--steps_per_epoch", default=10, type=int, help="length of the synthetic loader (stands for len(data_loader))")
    return p


def effective_lr(args, world_size):
    """main_pretrain.py:239-243: only the base rate given -> lr = blr * (batch * accum * world) / 256."""
    eff = args.batch_size * args.accum_iter * world_size
    return (args.blr * eff / 256 if args.lr is None else args.lr), eff


def main(args):
    from passl_b200 import models
    from passl_b200.core import ParamStore
    from passl_b200.distributed import get_rank, get_world_size, grad_sync, param_sync
    from passl_b200.optimizer import AdamW
    from passl_b200.optimizer.lr import MAEHalfCycleCosine
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(args.seed + get_rank())
    model = getattr(models, args.model)(norm_pix_loss=args.norm_pix_loss, img_size=args.input_size).to(dev)
    store = ParamStore(model)
    param_sync(store)
    lr, eff = effective_lr(args, get_world_size())
    if get_rank() == 0:
        print("base lr: %.2e\nactual lr: %.2e\neffective batch size: %d" % (lr * 256 / eff, lr, eff), flush=True)
    opt = AdamW(store, lr=lr, beta1=0.9, beta2=0.95, weight_decay=args.weight_decay, one_dim_no_decay=True)
    sched = MAEHalfCycleCosine(lr, args.min_lr, args.warmup_epochs, args.epochs, args.steps_per_epoch)
    g = torch.Generator(device=dev).manual_seed(1234 + get_rank())
    samples = torch.randn(args.batch_size, 3, args.input_size, args.input_size, device=dev, generator=g)
    step = 0
    for epoch in range(args.epochs):
        tic = time.time()
        for it in range(args.steps_per_epoch):
            if args.max_train_step is not None and step >= args.max_train_step:
                if get_rank() == 0:
                    print("step(%d) >= max_train_step(%d), training stops early." % (step, args.max_train_step), flush=True)
                return step
            # engine_pretrain.py:48-85: the rate moves and the optimizer steps every accum_iter iterations; the loss of each
            # iteration is divided by accum_iter (folded into the optimizer's gradient multiplier), gradients add up in between
            if it % args.accum_iter == 0:
                opt.set_lr(sched.lr_at(step))                     # lr_sched.adjust_learning_rate(it / len + epoch)
            loss, _, _ = model(samples, mask_ratio=args.mask_ratio)
            loss.backward()
            step += 1
            if (it + 1) % args.accum_iter == 0:
                grad_sync(store)
                opt.grad_scale = 1.0 / (get_world_size() * args.accum_iter)
                opt.step()
                opt.clear_grad()
            if (it + 1) % args.print_freq == 0 or it + 1 == args.steps_per_epoch:
                value = float(loss.detach())
                if not math.isfinite(value):
                    print("Loss is {}, stopping training".format(value), flush=True)
                    sys.exit(1)
                if get_rank() == 0:
                    ips = args.batch_size * get_world_size() * (it + 1) / max(time.time() - tic, 1e-9)
                    print("Epoch: [%d]  [%d/%d]  lr: %.6f  loss: %.4f  ips: %.1f images/sec" % (epoch, it + 1, args.steps_per_epoch, opt.get_lr(),
                                                                                             value, ips), flush=True)
    return step


if __name__ == "__main__":
    main(get_args_parser().parse_args())
