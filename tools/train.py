#!/usr/bin/env python
"""passl_b200 train entry (tools_v110/train.py / tools/train.py surface): `python tools/train.py -c configs/moco/moco_v2_r50.yaml
-o key=value`; multi-GPU through torch.distributed.run (one rank per GPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passl_b200.engine.trainer import Trainer  # noqa: E402
from passl_b200.utils import config as cfg_util  # noqa: E402


def main():
    args = cfg_util.parse_args()
    cfg = cfg_util.get_config(args.config, overrides=args.override)
    cfg["profiler_options"] = args.profiler_options
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    if "Global" in cfg:                  # v2.5 schema (Global / Model / Optimizer / LRScheduler / DataLoader): tools/train.py:25-32
        from passl_b200.engine.engine import Engine
        engine = Engine(cfg, mode="train")
        if args.resume:
            engine.resume(args.resume)
        engine.train()
        return
    trainer = Trainer(cfg)
    if args.resume:                      # tools_v110/train.py:35-40: continue a run, or start from weights only
        trainer.resume(args.resume)
    elif args.load:
        trainer.load(args.load)
    trainer.train()


if __name__ == "__main__":
    main()
