#!/usr/bin/env python
"""Benchmark of the SRDiff reverse-diffusion hot path (BASELINE.json metric: SR images/sec,
4x, 160x160 outputs, 100 timesteps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full `GaussianDiffusion.sample()` pass over one batch of synthetic LR patches
(RRDB encoder once + T denoiser steps + res2img).  Workload at every N: BASELINE.json configs[1]
per GPU (batch 64 of 40x40 -> 160x160, T=100, bf16), i.e. weak scaling; for N > 1 the timed
region also contains the final NCCL all-gather of the outputs (configs[3]'s exchange step).

`--impl reference` times the CPU implementation of the same path (the oracle port of the
reference's PyTorch code -- the Python reference cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# exact conv work per 160x160 image (SURVEY 8d / BASELINE.md 2): 2*MAC, no padding counted
GFLOP_UNET_STEP = 33.4795      # per image-step, cond_proj hoisted (what this build executes)
GFLOP_ONCE = 6.9971            # RRDB encoder + cond_proj, once per image


def load_hparams():
    hp = json.load(open(os.path.join(ROOT, "tests", "golden", "hparams.json")))
    hp["unet_dim_mults"] = tuple(hp["unet_dim_mults"])
    return hp


def unet_pass_traffic(B, lr):
    """DRAM bytes of one Unet pass (dram__bytes_read.sum + dram__bytes_write.sum over its conv launches) from
    the committed ncu launch list of this very command; only valid for the configuration it was captured on."""
    p = os.path.join(ROOT, "profiles", "r01", "unet_pass_traffic.json")
    if B == 64 and lr == 40 and os.path.isfile(p):
        d = json.load(open(p))
        return d["dram_bytes_read"] + d["dram_bytes_write"], "profiles/r01/unet_pass_traffic.json"
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", 1400.5), "tflops_burst": d.get("bf16_tflops", 1654.3),
                "hbm_gbs": d.get("hbm_gbs", 6473.9), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # "under load": drop idle samples (clock below half of the median of the top half)
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's PyTorch implementation, bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_reference_rate(hp, lr, T, batch, n_unet_steps, threads=None):
    """images/s of the CPU path, extrapolated to T steps from rrdb once + n_unet_steps p_sample steps
    at `batch` images (per-step cost is constant: SURVEY 8d, 100 x 0.148 s = 14.85 s measured)."""
    import torch
    from oracle import srdiff_oracle as O
    import dgm_img_super_resolution_b200.model as M
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    M.hparams.clear()
    M.hparams.update(hp)
    torch.manual_seed(1234)
    nf = hp["rrdb_num_feat"]
    rrdb = M.RRDBNet(3, 3, nf, hp["rrdb_num_block"], nf // 2)
    unet = M.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
    diff = M.GaussianDiffusion(unet, rrdb, timesteps=T)
    sd = {k: v.clone() for k, v in diff.state_dict().items()}
    sched = O.make_schedule(T, hp)
    g = torch.Generator().manual_seed(1)
    s = hp["sr_scale"]
    img_lr = torch.rand((batch, 3, lr, lr), generator=g) * 2 - 1
    x = torch.randn((batch, 3, lr * s, lr * s), generator=g)
    with torch.no_grad():
        O.rrdb_forward(sd, img_lr[:1], hp)                      # warm-up (thread pool, oneDNN primitives)
        t0 = time.perf_counter()
        _, feas = O.rrdb_forward(sd, img_lr, hp)
        t_rrdb = time.perf_counter() - t0
        t0 = time.perf_counter()
        for k in range(n_unet_steps):
            t = torch.full((batch,), T - 1 - k, dtype=torch.long)
            eps = O.unet_forward(sd, x, t, feas, hp)
            x, _ = O.p_sample_update(sched, x, t, eps, torch.randn(x.shape, generator=g))
        t_step = (time.perf_counter() - t0) / n_unet_steps
    total = t_rrdb + T * t_step
    return {"value": batch / total, "t_rrdb_s": t_rrdb, "t_step_s": t_step, "cores": torch.get_num_threads(),
            "sample": "oracle port, fp32, batch %d of %dx%d->%dx%d: RRDB once + %d p_sample steps timed, "
                      "extrapolated to T=%d (t_rrdb + T*t_step)" % (batch, lr, lr, lr * s, lr * s, n_unet_steps, T)}


def run_reference_arm(args, hp):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(hp, args.lr, args.timesteps, args.cpu_batch, args.cpu_unet_steps)
        if i >= args.warmup:
            vals.append(r["value"])
        last = r
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "SR images/sec (4x, 160x160, 100 steps)", "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * args.batch / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SRDiff 4x batch %d of %dx%d->%dx%d patches, %d timesteps (configs[1]); each step "
                               "is a bounded CPU sample" % (args.batch, args.lr, args.lr, 4 * args.lr, 4 * args.lr,
                                                            args.timesteps)},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": last["cores"], "kind": "port",
                         "sample": last["sample"]},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, hp):
    import torch
    import torch.distributed as dist
    import dgm_img_super_resolution_b200.model as M
    from dgm_img_super_resolution_b200 import engine as E, _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.lib()

    B, lr, T, s = args.batch, args.lr, args.timesteps, hp["sr_scale"]
    H = lr * s
    M.hparams.clear()
    M.hparams.update(hp)
    torch.manual_seed(1234)
    nf = hp["rrdb_num_feat"]
    rrdb = M.RRDBNet(3, 3, nf, hp["rrdb_num_block"], nf // 2)
    unet = M.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
    diff = M.GaussianDiffusion(unet, rrdb, timesteps=T).to(dev).eval()
    diff.sample_tqdm = False
    M.runtime.precision, M.runtime.noise, M.runtime.graph = args.precision, "philox", True
    M.runtime.image_offset = rank * B                       # shard-invariant Philox streams

    g = torch.Generator().manual_seed(1234 + rank)
    img_lr_h = (torch.rand((B, 3, lr, lr), generator=g) * 2 - 1).pin_memory()
    img_lr_up_h = torch.nn.functional.interpolate(img_lr_h, scale_factor=s, mode="bicubic",
                                                  align_corners=True).contiguous().pin_memory()
    out_h = torch.empty((B, 3, H, H)).pin_memory()
    img_lr, img_lr_up = img_lr_h.to(dev), img_lr_up_h.to(dev)
    shape = tuple(img_lr_up.shape)
    gather_buf = torch.empty((world * B, 3, H, H), device=dev) if world > 1 else None

    def one_pass():
        img, _ = diff.sample(img_lr, img_lr_up, shape)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, img)
        return img

    def one_pass_e2e():
        a = img_lr_h.to(dev, non_blocking=True)
        b = img_lr_up_h.to(dev, non_blocking=True)
        img, _ = diff.sample(a, b, shape)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, img)
        out_h.copy_(img, non_blocking=True)                 # stream-ordered D2H; synchronised by timed()
        return out_h

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 1)):
        one_pass()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(one_pass, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(1):
        one_pass_e2e()
    ms_e2e = timed(one_pass_e2e, args.steps)
    value = world * B * args.steps / (ms / 1000.0)
    value_e2e = world * B * args.steps / (ms_e2e / 1000.0)

    # ---- roofline of the dominant kernel family (tcgen05 conv): one Unet pass = n_conv launches
    plan = diff._plan(B, lr, lr, dev)
    n_ops, n_tc = plan.unet_prog.bld.count_kernels()
    temb_ptr = plan.temb.data_ptr()
    for _ in range(2):
        plan.unet_prog.run(temb_ptr)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.unet_prog.run(temb_ptr)
    e1.record()
    torch.cuda.synchronize()
    ms_unet = e0.elapsed_time(e1) / reps
    peaks = measured_peaks()
    traffic, traffic_src = unet_pass_traffic(B, lr)
    flops_per_pass = GFLOP_UNET_STEP * (H * H / 25600.0) * B * 1e9
    achieved = flops_per_pass / (ms_unet / 1000.0) / 1e12

    # ---- p_sample kernel: HBM roofline (12 B/element with in-kernel Philox + NHWC copy).  The launches are
    # replayed from a CUDA graph, as in the sampler: launched one by one from Python the ~10 us kernel would
    # be timed at the host's launch rate.
    eps = plan.unet_prog.eps
    n_ps = 20

    def ps_launches():
        for k in range(n_ps):
            E.p_sample_step(plan.x, eps, None, plan.coef, plan.t_idx[5], plan.x, True, 1, k, 0,
                            nhwc_out=plan.unet_prog.xin, act_dtype=plan.prec.dtype)
    ps_launches()
    torch.cuda.synchronize()
    g_ps = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_ps):
        ps_launches()
    g_ps.replay()
    torch.cuda.synchronize()
    e0.record()
    g_ps.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_ps = e0.elapsed_time(e1) / n_ps
    ps_bytes = plan.x.numel() * 12.0
    ps_gbs = ps_bytes / (ms_ps / 1000.0) / 1e9

    layers = None
    if args.layers and rank == 0:
        layers = per_layer_timings(plan, torch, E, L)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(layers, open(os.path.join(ROOT, "gpurun_out", "layers.json"), "w"), indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference_rate(hp, lr, T, args.cpu_batch, args.cpu_unet_steps)

    if rank == 0:
        line = {
            "metric": "SR images/sec (4x, 160x160, 100 steps)", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": "SRDiff 4x batch %d of %dx%d->%dx%d patches per GPU, %d timesteps, random-init "
                                   "weights from srdiff/config.yaml (configs[1])" % (B, lr, lr, H, H, T),
                       "global_batch": world * B, "parallelism": "batch-sharded x%d, final all_gather" % world,
                       "noise": "in-kernel Philox", "cuda_graph": True,
                       "l2": "working set per Unet pass (%.1f GB of activations) is far larger than the 126 MB L2"
                             % (78.3e6 * B * (H * H / 25600.0) / 1e9)},
            "clocks": clocks,
            "e2e": {"value": value_e2e, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(img_lr_h.numel() * 4 + img_lr_up_h.numel() * 4),
                    "d2h_bytes_per_step": int(out_h.numel() * 4)},
            "gpu_launches": int(args.steps * plan.n_launches()),
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (all %d tcgen05 conv launches of one Unet "
                                                      "pass; %d launches total per pass)" % (n_tc, n_ops),
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"], "traffic": traffic, "traffic_unit": "bytes per Unet pass",
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_pass": 78.3e6 * B * (H * H / 25600.0),
                         "peak_source": peaks["source"] + ", sustained bf16",
                         "flops_per_pass": flops_per_pass, "ms_per_pass": ms_unet},
            "roofline_p_sample": {"bound": "hbm", "achieved": ps_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                  "frac": ps_gbs / peaks["hbm_gbs"], "bytes_per_launch": ps_bytes,
                                  "us_per_launch": ms_ps * 1000.0},
            "whole_job_tensor_frac": (GFLOP_UNET_STEP * T + GFLOP_ONCE) * (H * H / 25600.0) * 1e9 * value
            / world / 1e12 / peaks["tflops"],
            "cpu_baseline": None if cpu is None else {"value": cpu["value"], "unit": "images/s",
                                                      "cores": cpu["cores"], "kind": "port",
                                                      "sample": cpu["sample"]},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def per_layer_timings(plan, torch, E, L):
    """CUDA-event time of every launch of one Unet pass (each op repeated 5x back to back)."""
    out = []
    ctx = E.RunCtx(plan.temb.data_ptr())
    for op in plan.unet_prog.bld.ops:
        if not isinstance(op, E.ConvOp):
            continue
        d = op.desc
        for _ in range(2):
            op.run(ctx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            op.run(ctx)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000.0 / 5
        cin = sum(d.src[i].cin for i in range(d.n_src))
        taps = d.nky * d.nkx
        if d.src[0].flags & 1:          # SRD_SRC_WIN8: three K=64 blocks stand for a 3x3 conv over the 3 real channels
            cin, taps = 3, 9
        flops = 2.0 * d.B * (d.H_out * d.W_out) * d.N * cin * taps
        out.append({"cin": cin, "N": d.N, "H_out": d.H_out, "stride": d.stride, "up": d.up, "taps": taps,
                    "kernel": op.kernel, "us": us, "tflops": flops / us / 1e6})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU (configs[1]: 64)")
    ap.add_argument("--lr", type=int, default=40, help="LR patch side (configs[1]: 40 -> 160)")
    ap.add_argument("--timesteps", type=int, default=100)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16_direct"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--cpu-unet-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--layers", action="store_true", help="also dump per-layer timings to gpurun_out/layers.json")
    args = ap.parse_args()
    hp = load_hparams()
    if args.impl == "reference":
        run_reference_arm(args, hp)
    else:
        run_ours(args, hp)


if __name__ == "__main__":
    main()
