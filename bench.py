#!/usr/bin/env python
"""bench.py — images/sec of the PASSL self-supervised hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] — SimCLR ResNet-50, bf16 tensor-core math, 224^2, 512 samples x 2
views per GPU (global batch 4096 at 8 GPUs, weak scaling), all-gathered negatives, NT-Xent + CO2 head, LARS step.
A "step" = forward (1024 images through ResNet-50 + fc3 neck) + loss + backward + gradient all-reduce + optimizer.

  python bench.py --gpus N --steps K --warmup W           (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                    the reference math on the host CPU cores (oracle port)

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: same step fed from pinned host memory with a
device->host read of the loss every step; `roofline`: the dominant kernel (tcgen05 implicit-GEMM) against the measured
bf16 peak, plus the fused InfoNCE kernel against the measured HBM peak (the second half of BASELINE's metric).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = int(os.environ.get("PASSL_B200_BENCH_BATCH", "512"))
IMG = 224
METRIC = "images/sec (224^2) at 1/2/4/8 B200; fused InfoNCE HBM GB/s vs roofline"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                    reasons=sorted(reasons), samples=len(self.samples))


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the same training iteration on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step_time(steps, warmup, sample):
    import torch
    from oracle import simclr_step as S          # bench.py may execute oracle/ only here (cpu_baseline / --impl reference)
    # torch-CPU convolutions stop scaling (and thrash) far below the core count of the GPU host: cap the intra-op pool
    cores = min(os.cpu_count() or 1, int(os.environ.get("PASSL_B200_CPU_THREADS", "32")))
    torch.set_num_threads(cores)
    p = S.init_params(0)
    vel = {}
    g = torch.Generator().manual_seed(1234)
    a = torch.randn(sample, 3, IMG, IMG, generator=g)
    b = torch.randn(sample, 3, IMG, IMG, generator=g)
    for _ in range(warmup):
        S.train_step(p, vel, a, b, lr=1e-3)
    t0 = time.perf_counter()
    for _ in range(steps):
        S.train_step(p, vel, a, b, lr=1e-3)
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    ips, dt, cores = cpu_reference_step_time(args.steps, min(args.warmup, 1), sample)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "simclr_r50_224_2views (BASELINE configs[1]); CPU oracle port of the reference math",
                       "per_step_samples": sample},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "%d samples x 2 views of 3x224x224 per step (full fwd+bwd+LARS)" % sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from passl_b200 import _lib, kernels as K
    from passl_b200.core import ParamStore
    from passl_b200.distributed import grad_sync, param_sync
    from passl_b200.modeling import build_model
    from passl_b200.optimizer import LarsMomentumOptimizer
    lib = _lib.load()
    pk = peaks()

    B = PER_GPU_BATCH
    torch.manual_seed(0)
    model = build_model(dict(name="SimCLR",
                             backbone=dict(name="ResNet", depth=50, with_pool=True),
                             neck=dict(name="NonLinearNeckfc3", in_channels=2048, hid_channels=2048, out_channels=128,
                                       with_avg_pool=False),
                             head=dict(name="SimCLRContrastiveHead", temperature=0.1, multi_rank=True))).to(dev)
    store = ParamStore(model.encoder)
    param_sync(store)
    base_lr = 0.075 * (B * world) ** 0.5            # learning_rate_scaling: sqrt  (configs/simclr/simclr_r50_IM.yaml)
    warm_steps = 10 * 1281167 // (B * world)
    opt = LarsMomentumOptimizer(store, lr=0.0)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    view_a = torch.randn(B, 3, IMG, IMG, device=dev, generator=gen)
    view_b = torch.randn(B, 3, IMG, IMG, device=dev, generator=gen)
    host_a = torch.randn(B, 3, IMG, IMG).pin_memory()
    host_b = torch.randn(B, 3, IMG, IMG).pin_memory()
    stage_a, stage_b = torch.empty_like(view_a), torch.empty_like(view_b)
    it = [0]

    def step(a, b):
        opt.set_lr(base_lr * min(1.0, (it[0] + 1) / warm_steps))
        opt.clear_grad()
        out = model(a, b)
        out["loss"].backward()
        grad_sync(store)
        opt.step()
        it[0] += 1
        return out["loss"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            loss = fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), loss

    for _ in range(args.warmup):
        step(view_a, view_b)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.passl_b200_launch_count()
    ms, loss = timed(lambda: step(view_a, view_b), args.steps)
    launches = lib.passl_b200_launch_count() - launches0
    sampler.stop_flag = True
    loss_val = float(loss.item())
    value = B * world * args.steps / (ms / 1e3)

    # ---- e2e: inputs from pinned host memory every step + D2H read of the loss -------------------------------------
    # The user-facing loop (engine/trainer.py IterLoader with prefetch) double-buffers the input: the H2D copy of step t+1 runs
    # on a copy stream under the compute of step t.  Every timed step still performs one full H2D copy of its inputs (K copies
    # inside the timed region for K steps) and one D2H read of its loss.
    copy_stream = torch.cuda.Stream()
    bufs = [(stage_a, stage_b), (torch.empty_like(view_a), torch.empty_like(view_b))]
    ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
    ev_free = [torch.cuda.Event(), torch.cuda.Event()]
    kk = [0]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[i])
            bufs[i][0].copy_(host_a, non_blocking=True)
            bufs[i][1].copy_(host_b, non_blocking=True)
            ev_ready[i].record(copy_stream)

    def e2e_step():
        i = kk[0] & 1
        prefetch(i ^ 1)                                 # next step's inputs: overlaps this step's compute
        torch.cuda.current_stream().wait_event(ev_ready[i])
        l = step(*bufs[i])
        ev_free[i].record()
        kk[0] += 1
        return l.item()                                 # D2H read of the step result (sync), like loop.py:86
    for e_ in ev_free:
        e_.record()
    prefetch(0)
    for _ in range(2):
        e2e_step()
    ms_e2e, _ = timed(e2e_step, args.steps)
    e2e_value = B * world * args.steps / (ms_e2e / 1e3)
    h2d = 2 * B * 3 * IMG * IMG * 4

    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "simclr_r50_224_2views_bs%d_per_gpu (BASELINE configs[1]: global bs %d at 8 GPUs)" % (B, 8 * B),
                       "backbone": "ResNet-50 v1.5 (stem max-pool; reference ResNetsimclr variant available as stem_maxpool=False)",
                       "global_batch": B * world, "views": 2, "head": "NT-Xent+CO2, all-gathered negatives", "optimizer": "LARS",
                       "parallelism": "dp%d" % world, "l2_policy": "inputs (617 MB/step) and activations exceed the 126 MB L2",
                       "images_counted": "samples per step (each sample = two 224^2 views = 2 backbone passes)",
                       "final_loss": loss_val},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches)}

    # ---- roofline of the dominant kernel: instrumented (untimed) step with CUDA events around every tcgen05 launch.
    #      Every rank runs the step (it contains collectives); only rank 0 reports.
    if True:
        if rank == 0:
            line["clocks"] = sampler.summary()
        rec = []
        orig = {n: getattr(K, n) for n in ("gemm", "conv2d_fwd", "conv2d_dgrad", "conv2d_wgrad")}

        def cost_of(name, a, kw, res):
            """(algorithmic FLOPs, algorithmic HBM bytes) of one launch: operands read once, output written once."""
            def nb(t):
                return 0 if t is None else t.numel() * t.element_size()
            extra = nb(kw.get("residual")) + nb(kw.get("aux")) + nb(kw.get("preact_out"))
            if name == "gemm":
                A, Bm = a[0], a[1]
                Kd = A.shape[0] if kw.get("a_t") else A.shape[1]
                acc = nb(res) if kw.get("accumulate") else 0
                return 2.0 * res.shape[0] * res.shape[1] * Kd, nb(A) + nb(Bm) + nb(res) + acc + extra
            if name == "conv2d_fwd":
                x, w = a[0], a[1]
                return 2.0 * res.numel() * w.shape[1] * w.shape[2] * w.shape[3], nb(x) + nb(w) + nb(res) + extra
            if name == "conv2d_dgrad":
                dy, w = a[0], a[1]
                acc = nb(res) if kw.get("accumulate") else 0
                return 2.0 * dy.numel() * w.shape[1] * w.shape[2] * w.shape[3], nb(dy) + nb(w) + nb(res) + acc
            x, dy, ws = a[0], a[1], a[2]
            return 2.0 * dy.numel() * ws[1] * ws[2] * ws[3], nb(x) + nb(dy) + 2 * nb(res)

        def wrap(name):
            def f(*a, **kw):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = orig[name](*a, **kw)
                e.record()
                rec.append((s, e) + cost_of(name, a, kw, r))
                return r
            return f
        from passl_b200.core import streams
        side_was = streams.ENABLED
        streams.ENABLED = False                         # serial launches: per-launch durations are not inflated by overlap
        for n in orig:
            setattr(K, n, wrap(n))
        step(view_a, view_b)
        torch.cuda.synchronize()
        for n, f in orig.items():
            setattr(K, n, f)
        streams.ENABLED = side_was
    if rank == 0:
        # Every tcgen05 launch is classed by ITS binding roofline: t_tensor = FLOPs / bf16 peak, t_hbm = bytes / HBM peak.
        # ResNet-50's 1x1 convolutions at 56^2 / 28^2 have < 218 FLOP/B and are HBM-bound even on tensor cores.
        pk_tf, pk_bw = pk["bf16_sustained"] * 1e12, pk["hbm_gbs"] * 1e9
        cls = {"tensor": [0.0, 0.0, 0.0, 0], "hbm": [0.0, 0.0, 0.0, 0]}      # ms, flops, bytes, launches
        t_bind = 0.0
        for s_, e_, fl, by in rec:
            dt = s_.elapsed_time(e_)
            k = "tensor" if fl / pk_tf >= by / pk_bw else "hbm"
            c = cls[k]
            c[0] += dt; c[1] += fl; c[2] += by; c[3] += 1
            t_bind += max(fl / pk_tf, by / pk_bw) * 1e3
        tc_ms = cls["tensor"][0] + cls["hbm"][0]
        tc_flops = cls["tensor"][1] + cls["hbm"][1]

        def roof(k):
            ms_k, fl, by, nl = cls[k]
            if k == "tensor":
                ach, peak, unit = fl / (ms_k / 1e3) / 1e12 if ms_k else 0.0, pk["bf16_sustained"], "TFLOP/s"
            else:
                ach, peak, unit = by / (ms_k / 1e3) / 1e9 if ms_k else 0.0, pk["hbm_gbs"], "GB/s"
            # traffic: dram__bytes_read.sum + dram__bytes_write.sum of ONE representative launch of the class from the committed
            # `ncu --set full` capture (profiles/r01_ncu_c3_full_summary.txt): hbm class = 1x1 conv 64->256 at 56^2, B=256
            # (M=802816, N=256, K=64, fused BN statistics): 104.1 MB read + 357.7 MB written by kernel end vs 102.8 + 411.0 MB
            # algorithmic (the rest of the output is still dirty in L2) — no re-reads.  tensor class: 3x3 conv 64->64 at 56^2, B=128
            # (profiles/r01_ncu_conv_full_summary.txt): 51.5 MB read vs 51.4 MB algorithmic input.
            traffic = {"hbm": {"bytes": 461.8e6, "algorithmic_bytes": 513.8e6, "launch": "gemm<256,64,0,0> M=802816 N=256 K=64 +BN stats",
                               "source": "profiles/r01_ncu_c3_full_summary.txt"},
                       "tensor": {"bytes": 57.2e6, "algorithmic_bytes": 102.8e6, "launch": "conv3x3 64->64 56^2 B=128 fwd",
                                  "source": "profiles/r01_ncu_conv_full_summary.txt"}}[k]
            return {"kernel": "gemm_tcgen05_kernel (implicit-GEMM conv fwd/dgrad/wgrad + linears), %s-bound launches" % k,
                    "bound": k, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "traffic": traffic["bytes"],
                    "traffic_detail": traffic,
                    "peak_source": pk["src"], "launches": nl, "share_of_step": ms_k / (ms / args.steps)}
        dom = "hbm" if cls["hbm"][0] >= cls["tensor"][0] else "tensor"
        other = "tensor" if dom == "hbm" else "hbm"
        line["roofline"] = roof(dom)
        line["roofline"].update({
            "all_launches": len(rec), "all_share_of_step": tc_ms / (ms / args.steps), "flops_per_step": tc_flops,
            "all_tflops": tc_flops / (tc_ms / 1e3) / 1e12,
            "frac_of_binding_roofline_all_launches": t_bind / tc_ms,
            "note": "one instrumented step (side stream off), CUDA events around each of the %d tcgen05 launches; each launch is "
                    "classed by its binding roofline (algorithmic FLOPs / bf16 peak vs algorithmic bytes / HBM peak); this object "
                    "is the class with the larger time share, roofline_other the rest; traffic from profiles/ (ncu --set full)"
                    % len(rec)})
        line["roofline_other"] = roof(other)
        # ---- fused InfoNCE (MoCo C3 shape) against the HBM roofline: CUDA-graph replay, events on the capture stream --------
        N, D, Kq, T = 256, 128, 65536, 0.2
        NQ = 8                                            # 8 distinct queues (8 x 16.8 MB = 134 MB > 126 MB L2): none is L2 resident
        q = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=1)
        kpos = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=1)
        queues = [torch.nn.functional.normalize(torch.randn(Kq, D, device=dev), dim=1).bfloat16() for _ in range(NQ)]
        qb = q.bfloat16()
        K.infonce_tc_fwd(qb, queues[0], pos=kpos, scale=1 / T)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        gph = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            K.infonce_tc_fwd(qb, queues[0], pos=kpos, scale=1 / T)
            with torch.cuda.graph(gph, stream=st):
                for qq in queues:                         # one graph = NQ back-to-back forward calls over different queues
                    K.infonce_tc_fwd(qb, qq, pos=kpos, scale=1 / T)
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.zero_()                              # L2 flush between replays
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            gph.replay()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / NQ)
        ts.sort()
        t_med = ts[len(ts) // 2]
        alg_bytes = (2 * N * D + D * Kq) * 2 + 4 * N
        gbs = alg_bytes / (t_med / 1e3) / 1e9
        line["roofline_infonce"] = {"kernel": "infonce_target_kernel + infonce_tc_fwd_kernel<2> + simce_finalize_kernel (MoCo C3: N=256, K=65536, D=128, bf16)",
                                    "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                                    # dram__bytes_read.sum + write of infonce_tc_fwd_kernel, profiles/r01_ncu_full_infonce_run8.txt
                                    "traffic": 17.04e6, "algorithmic_bytes": alg_bytes, "us_per_launch": t_med * 1e3,
                                    "peak_source": pk["src"] + " (burst)",
                                    "method": "CUDA graph of %d forward calls over %d different queues (working set %d MB > L2, flushed "
                                              "between replays), device time / %d" % (NQ, NQ, NQ * D * Kq * 2 >> 20, NQ)}
        # ---- cpu_baseline (N=1 only): bounded sample of the same iteration on the host cores -------------------------------
        if world == 1 and not args.no_cpu_baseline:
            ips, dt, cores = cpu_reference_step_time(1, 1, 8)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "1 timed step of 8 samples x 2 views (fwd+bwd+LARS) after 1 warm-up; torch-CPU fp32 oracle port"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
