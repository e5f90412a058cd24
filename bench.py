#!/usr/bin/env python
"""Benchmark of the SRDiff reverse-diffusion hot path (BASELINE.json metric: SR images/sec,
4x, 160x160 outputs, 100 timesteps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one full `GaussianDiffusion.sample()` pass over one batch of synthetic LR inputs
(RRDB encoder once + T denoiser steps + res2img).  Workloads (BASELINE.json `configs`):

    c1        configs[1]: batch 64 of 40x40 -> 160x160 per GPU, T = 100 (default; weak scaling over --gpus)
    c2_whole  configs[2]: one 256x256 -> 1024x1024 image as a whole (what st_app1.py feeds the model), T = 100
    c2_tiled  configs[2]: the same image cut into 16 tiles of 64x64 sampled as one batch and merged (pipeline.enhance)
    c3        configs[3]: GLOBAL batch 1024 of 160x160 outputs sharded over the GPUs (strong scaling), gathered
    c4_T25 | c4_T50 | c4_T100 | c4_T1000   configs[4]: batch 128 of 64x64 -> 256x256, timestep sweep

For N > 1 every workload goes through the product's own `dist.sample_sharded` (batch sharding + the final NCCL
gather of the outputs, the path's only collective) and the line carries `shard_invariance`: after the timed
region rank 0 re-samples another rank's slice and a ragged batch and compares them bit for bit with the gathered rows.

`--impl reference` times the reference's own CPU implementation of the same path on the host cores: the unmodified
`srdiff/model.py` (copied to the git-ignored `oracle/_ref/` by `oracle/build_ref.py`; kind "reference") when it is
there, else the oracle port (kind "port").  Each step is a bounded sample: the full batch through ONE p_sample
step (+ the encoder, timed once), extrapolated to T steps (per-step cost is constant: SURVEY 8d).
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
MB_UNET_STEP = 78.3            # minimum bf16 NHWC activation traffic per image-step at 160x160
METRIC = "SR images/sec (4x, 160x160, 100 steps)"

WORKLOADS = {
    # name: per-GPU batch (or global for c3), LR side, T, scaling, description
    "c1": dict(batch=64, lr=40, T=100, scaling="weak",
               desc="SRDiff 4x batch 64 of 40x40->160x160 patches per GPU, 100 timesteps (configs[1])"),
    "c2_whole": dict(batch=1, lr=256, T=100, scaling="weak",
                     desc="SRDiff 4x one 256x256->1024x1024 image sampled whole, 100 timesteps (configs[2])"),
    "c2_tiled": dict(batch=1, lr=256, T=100, scaling="weak", tile=64,
                     desc="SRDiff 4x one 256x256->1024x1024 image as 16 tiles of 64x64 in one batch, merged, "
                          "100 timesteps (configs[2])"),
    "c3": dict(batch=1024, lr=40, T=100, scaling="strong", micro=128,
               desc="SRDiff 4x GLOBAL batch 1024 of 40x40->160x160 sharded over the GPUs with output gather, "
                    "100 timesteps (configs[3])"),
}
for _T in (25, 50, 100, 1000):
    WORKLOADS["c4_T%d" % _T] = dict(batch=128, lr=64, T=_T, scaling="weak",
                                    desc="SRDiff 4x batch 128 of 64x64->256x256 per GPU, %d timesteps (configs[4])" % _T)


def load_hparams():
    hp = json.load(open(os.path.join(ROOT, "tests", "golden", "hparams.json")))
    hp["unet_dim_mults"] = tuple(hp["unet_dim_mults"])
    return hp


def unet_pass_traffic(B, lr):
    """DRAM bytes of one Unet pass (dram__bytes_read.sum + dram__bytes_write.sum over its conv launches) from the
    newest committed ncu launch list of this very command; only valid for the configuration it was captured on."""
    for rnd in ("r02", "r01"):
        p = os.path.join(ROOT, "profiles", rnd, "unet_pass_traffic.json")
        if B == 64 and lr == 40 and os.path.isfile(p):
            d = json.load(open(p))
            return d["dram_bytes_read"] + d["dram_bytes_write"], "profiles/%s/unet_pass_traffic.json" % rnd
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
        self.index, self.rows, self.proc, self.limit = index, [], None, None

    def _limit(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=power.limit",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            return float(out.strip().splitlines()[0])
        except Exception:
            return None

    def start(self):
        self.limit = self._limit()
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
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            try:
                pw.append(float(r[2]))
            except ValueError:
                pass
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # "under load": drop idle samples (clock below half of the median of the top half)
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        pw_load = sorted(pw)[len(pw) // 2:] if pw else []
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(pw_load) if pw_load else None, "power_limit_w": self.limit}


def synth_inputs(torch, B, lr, s, seed):
    g = torch.Generator().manual_seed(seed)
    img_lr = torch.rand((B, 3, lr, lr), generator=g) * 2 - 1
    img_lr_up = torch.nn.functional.interpolate(img_lr, scale_factor=s, mode="bicubic", align_corners=True).contiguous()
    return img_lr, img_lr_up, g


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref) or the oracle port, bounded sample
# ---------------------------------------------------------------------------------------------
class CpuPath:
    """The path on the host cores: `kind` "reference" = the unmodified srdiff/model.py (oracle/_ref, placed there by
    oracle/build_ref.py in the build container; /root/reference itself does not exist on the GPU box), else "port" =
    oracle/srdiff_oracle.py.  Same seeded weights either way."""

    def __init__(self, hp, T, threads=None, force_port=False):
        import torch
        self.torch = torch
        torch.set_num_threads(threads or len(os.sched_getaffinity(0)))   # torchrun exports OMP_NUM_THREADS=1
        self.hp, self.T = hp, T
        ref_root = os.path.join(ROOT, "oracle", "_ref")
        self.kind = "port"
        if not force_port and os.path.isfile(os.path.join(ref_root, "srdiff", "model.py")):
            os.environ["SRDIFF_REFERENCE_ROOT"] = ref_root
            from oracle import ref_loader
            ref_loader.REF_ROOT = ref_root
            self.ref, self.diff = ref_loader.build_reference(hp, timesteps=T, seed=1234)
            self.kind = "reference"
        else:
            from oracle import srdiff_oracle as O
            import dgm_img_super_resolution_b200.model as M
            M.hparams.clear()
            M.hparams.update(hp)
            torch.manual_seed(1234)
            nf = hp["rrdb_num_feat"]
            rrdb = M.RRDBNet(3, 3, nf, hp["rrdb_num_block"], nf // 2)
            unet = M.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
            diff = M.GaussianDiffusion(unet, rrdb, timesteps=T)
            self.O, self.sd = O, {k: v.clone() for k, v in diff.state_dict().items()}
            self.sched = O.make_schedule(T, hp)
        self.cores = torch.get_num_threads()

    def rrdb(self, img_lr):
        if self.kind == "reference":
            return self.diff.rrdb(img_lr, True)[1]
        return self.O.rrdb_forward(self.sd, img_lr, self.hp)[1]

    def p_sample(self, x, t_int, feas, img_lr_up):
        torch = self.torch
        t = torch.full((x.shape[0],), t_int, dtype=torch.long)
        if self.kind == "reference":
            return self.diff.p_sample(x, t, feas, img_lr_up)[0]
        eps = self.O.unet_forward(self.sd, x, t, feas, self.hp)
        return self.O.p_sample_update(self.sched, x, t, eps, torch.randn(x.shape))[0]

    def measure(self, batch, lr, n_unet_steps, cache):
        """images/s of `batch` images at T steps = batch / (t_rrdb + T * t_step); the encoder is timed once per
        process (`cache`), the p_sample step(s) on every call."""
        torch, s = self.torch, self.hp["sr_scale"]
        img_lr, img_lr_up, g = synth_inputs(torch, batch, lr, s, 1)
        x = torch.randn((batch, 3, lr * s, lr * s), generator=g)
        with torch.no_grad():
            if "feas" not in cache:
                self.rrdb(img_lr[:1])                       # warm-up (thread pool, oneDNN primitives)
                t0 = time.perf_counter()
                cache["feas"] = self.rrdb(img_lr)
                cache["t_rrdb"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            for k in range(n_unet_steps):
                x = self.p_sample(x, self.T - 1 - k, cache["feas"], img_lr_up)
            t_step = (time.perf_counter() - t0) / n_unet_steps
        total = cache["t_rrdb"] + self.T * t_step
        what = "the unmodified reference srdiff/model.py (oracle/_ref)" if self.kind == "reference" else \
            "oracle port of the reference's PyTorch code"
        return {"value": batch / total, "t_rrdb_s": cache["t_rrdb"], "t_step_s": t_step, "cores": self.cores,
                "kind": self.kind,
                "sample": "%s, fp32, %d CPU threads: the FULL batch of %d images %dx%d->%dx%d through the RRDB encoder "
                          "(timed once) + %d p_sample step(s) per bench step, extrapolated to T=%d "
                          "(t_rrdb + T*t_step; per-step cost is constant)" % (
                              what, self.cores, batch, lr, lr, lr * s, lr * s, n_unet_steps, self.T)}


def run_reference_arm(args, hp, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, lr, T = wl["eff_batch"], wl["lr"], wl["T"]
    cpu = CpuPath(hp, T, force_port=args.cpu_port)
    cache, vals, last = {}, [], None
    for i in range(args.warmup + args.steps):
        r = cpu.measure(min(B, args.cpu_batch), lr, args.cpu_unet_steps, cache)
        if i >= args.warmup:
            vals.append(r["value"])
        last = r
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * B / v, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"] + "; each step is a bounded CPU sample", "name": args.workload},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# cuDNN leg: the reference's ATen stream on the same B200 -- a reported baseline, outside the timed region
# ---------------------------------------------------------------------------------------------
def cudnn_baseline(torch, hp, dev, B, lr, T, n_steps=3):
    """images/s of the reference's own op stream (the oracle port's functional ATen calls = F.conv2d /
    conv_transpose2d / softplus-tanh Mish / cat ... -> cuDNN) executed eagerly on this GPU: fp32 with TF32 convs, and
    bf16 autocast + channels_last.  RRDB once + `n_steps` p_sample steps timed with CUDA events, extrapolated to T."""
    from oracle import srdiff_oracle as O
    import dgm_img_super_resolution_b200.model as M
    torch.manual_seed(1234)
    nf = hp["rrdb_num_feat"]
    rrdb = M.RRDBNet(3, 3, nf, hp["rrdb_num_block"], nf // 2)
    unet = M.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
    diff = M.GaussianDiffusion(unet, rrdb, timesteps=T)
    sd = {k: v.detach().to(dev) for k, v in diff.state_dict().items()}
    sched = {k: v.to(dev) for k, v in O.make_schedule(T, hp).items()}
    s = hp["sr_scale"]
    img_lr, img_lr_up, g = synth_inputs(torch, B, lr, s, 1)
    img_lr, img_lr_up = img_lr.to(dev), img_lr_up.to(dev)
    out = {}
    old_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ("tf32", "bf16_channels_last"):
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode != "tf32" else torch.autocast("cuda", enabled=False)
            x = torch.randn((B, 3, lr * s, lr * s), device=dev)
            if mode != "tf32":
                x = x.contiguous(memory_format=torch.channels_last)
            with torch.no_grad(), ctx:
                def step(x, k):
                    t = torch.full((B,), T - 1 - k, dtype=torch.long, device=dev)
                    eps = O.unet_forward(sd, x, t, feas, hp).float()
                    return O.p_sample_update(sched, x.float(), t, eps, torch.randn(x.shape, device=dev))[0]
                _, feas = O.rrdb_forward(sd, img_lr, hp)            # warm-up: cuDNN autotuning
                step(x, 0)
                torch.cuda.synchronize()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                _, feas = O.rrdb_forward(sd, img_lr, hp)
                e[1].record()
                for k in range(n_steps):
                    x = step(x, k)
                e[2].record()
                torch.cuda.synchronize()
            t_rrdb, t_step = e[0].elapsed_time(e[1]) / 1e3, e[1].elapsed_time(e[2]) / 1e3 / n_steps
            out[mode] = {"images_per_s": B / (t_rrdb + T * t_step), "ms_per_unet_step": 1e3 * t_step,
                         "ms_rrdb": 1e3 * t_rrdb}
            del feas
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old_tf32
    out["what"] = ("the reference's eager ATen/cuDNN op stream (oracle port's functional calls) on this GPU, batch %d: "
                   "RRDB once + %d p_sample steps (CUDA events), extrapolated to T=%d; cudnn.benchmark on" % (B, n_steps, T))
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, hp, wl):
    import torch
    import torch.distributed as dist
    import dgm_img_super_resolution_b200.model as M
    from dgm_img_super_resolution_b200 import engine as E, _lib as L, dist as D, pipeline as P

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.lib()

    lr, T, s = wl["lr"], wl["T"], hp["sr_scale"]
    H = lr * s
    strong = wl["scaling"] == "strong"
    # images each sample() call of one rank covers, and how many such calls one bench step makes
    if strong:
        G = wl["eff_batch"]                                   # global batch (c3: 1024)
        micro = min(wl.get("micro", G), max(1, G // world))   # per-rank micro-batch
        calls = max(1, G // (micro * world))
        Bcall_global = micro * world
    else:
        micro, calls = wl["eff_batch"], 1
        Bcall_global = micro * world
        G = Bcall_global
    M.hparams.clear()
    M.hparams.update(hp)
    torch.manual_seed(1234)
    nf = hp["rrdb_num_feat"]
    rrdb = M.RRDBNet(3, 3, nf, hp["rrdb_num_block"], nf // 2)
    unet = M.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
    diff = M.GaussianDiffusion(unet, rrdb, timesteps=T).to(dev).eval()
    diff.sample_tqdm = False
    M.runtime.precision, M.runtime.noise, M.runtime.graph = args.precision, "philox", True

    tile = wl.get("tile")
    # the GLOBAL batch of one sample_sharded call (same on every rank), in pinned host memory for the e2e leg
    img_lr_h, img_lr_up_h, _ = synth_inputs(torch, Bcall_global, lr, s, 1234)
    img_lr_h, img_lr_up_h = img_lr_h.pin_memory(), img_lr_up_h.pin_memory()
    out_h = torch.empty((Bcall_global, 3, H, H)).pin_memory()
    img_lr, img_lr_up = img_lr_h.to(dev), img_lr_up_h.to(dev)

    def sample_call(a, b):
        """One call of the public API for the global batch (a, b): returns the [Bcall_global, 3, H, H] images."""
        if tile:                                              # configs[2] tiled: split -> one batch of tiles -> merge
            return P.enhance(diff, a, tile=(tile, tile)), None
        if world > 1:
            return D.sample_sharded(diff, a, b)
        return diff.sample(a, b, tuple(b.shape))

    def one_pass():
        for _ in range(calls):
            img, _ = sample_call(img_lr, img_lr_up)
        return img

    # end to end: every rank uploads ITS OWN images from pinned host memory, rank 0 reads the gathered result back
    sh_start, sh_count = D.shard_range(Bcall_global, world, rank) if world > 1 else (0, Bcall_global)
    lr_sl, up_sl = img_lr_h[sh_start:sh_start + sh_count], img_lr_up_h[sh_start:sh_start + sh_count]

    def one_pass_e2e():
        for _ in range(calls):
            a = lr_sl.to(dev, non_blocking=True)
            b = up_sl.to(dev, non_blocking=True)
            if tile:
                P.enhance(diff, a, tile=(tile, tile)).cpu()
                continue
            if world > 1:
                img, _ = D.sample_sharded(diff, a, b, total=Bcall_global)
            else:
                img, _ = diff.sample(a, b, tuple(b.shape))
            if rank == 0:
                out_h.copy_(img, non_blocking=True)           # stream-ordered D2H; synchronised by timed()
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
    one_pass_e2e()
    ms_e2e = timed(one_pass_e2e, args.steps)
    value = G * args.steps / (ms / 1000.0)
    value_e2e = G * args.steps / (ms_e2e / 1000.0)

    # ---- multi-GPU: shard invariance of what the product path just produced
    shard_inv = None
    if world > 1 and not tile:
        M.runtime.seed_override = 20251018
        try:
            whole, _ = D.sample_sharded(diff, img_lr, img_lr_up)
            ok = True
            if rank == 0:                                     # rank 1's slice, re-sampled here with its image offset
                start, count = D.shard_range(Bcall_global, world, 1)
                M.runtime.image_offset = start
                one, _ = diff.sample(img_lr[start:start + count], img_lr_up[start:start + count],
                                     (count, 3, H, H))
                M.runtime.image_offset = 0
                ok = bool(torch.equal(one, whole[start:start + count]))
            # ragged global batch (not divisible by the world size), small shape: gather pads and trims
            Br = 3 * world + 1
            rl, ru, _ = synth_inputs(torch, Br, 8, s, 77)
            rl, ru = rl.to(dev), ru.to(dev)
            ragged, _ = D.sample_sharded(diff, rl, ru)
            if rank == 0:
                M.runtime.image_offset = 0
                ref, _ = diff.sample(rl, ru, tuple(ru.shape))  # the whole ragged batch on one GPU
                ok = ok and tuple(ragged.shape) == (Br, 3, 8 * s, 8 * s) and bool(torch.equal(ref, ragged))
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.broadcast(flag, src=0)
            shard_inv = bool(int(flag))
        finally:
            M.runtime.seed_override, M.runtime.image_offset = None, 0

    # ---- roofline of the dominant kernel family (tcgen05 conv): one Unet pass
    Bp = micro * (tile and (lr // tile) ** 2 or 1)
    lrp = tile or lr
    plan = diff._plan(Bp, lrp, lrp, dev)
    Hp = lrp * s
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
    traffic, traffic_src = unet_pass_traffic(Bp, lrp)
    px_scale = Hp * Hp / 25600.0
    flops_per_pass = GFLOP_UNET_STEP * px_scale * Bp * 1e9
    achieved = flops_per_pass / (ms_unet / 1000.0) / 1e12

    # ---- p_sample kernel: HBM roofline.  Honest bytes per element: x in, eps in, x' out (fp32) + the padded NHWC
    # bf16 copy of x' for the next step's first conv (8 channels per pixel incl. the ring).  Its 66 MB working set
    # would stay in the 126 MB L2 if the launch were simply repeated, while in the sampler a whole Unet pass (5 GB of
    # traffic) runs between two p_sample launches: every timed launch is preceded by a 512 MB read that evicts it
    # (clean lines, as the last conv's input leaves them), and timed with its own pair of events.
    eps = plan.unet_prog.eps
    n_ps = 20
    flush = torch.ones(512 << 20, dtype=torch.uint8, device=dev).view(torch.int32)
    sink = torch.zeros(1, dtype=torch.int64, device=dev)

    def ps_launches(with_ps):
        for k in range(n_ps):
            torch.sum(flush, dim=0, keepdim=True, out=sink)
            if with_ps:
                E.p_sample_step(plan.x, eps, None, plan.coef, plan.t_idx[5], plan.x, True, 1, k, 0,
                                nhwc_out=plan.unet_prog.xin, act_dtype=plan.prec.dtype)

    def timed_graph(with_ps):
        ps_launches(with_ps)
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            ps_launches(with_ps)
        g_.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0.record()
            g_.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)
    # graph replays (no host launch gaps); the flush alone is timed the same way and subtracted
    ms_ps = (timed_graph(True) - timed_graph(False)) / n_ps
    del flush
    ps_bytes = plan.x.numel() * 12.0 + plan.unet_prog.xin.t.numel() * plan.unet_prog.xin.t.element_size()
    ps_gbs = ps_bytes / (ms_ps / 1000.0) / 1e9

    layers = None
    if args.layers and rank == 0:
        layers = per_layer_timings(plan, torch, E, L)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(layers, open(os.path.join(ROOT, "gpurun_out", "layers.json"), "w"), indent=1)

    cpu = cud = None
    if rank == 0 and world == 1 and not args.no_cpu:
        path = CpuPath(hp, T, force_port=args.cpu_port)
        cpu = path.measure(min(wl["eff_batch"], args.cpu_batch), lr, args.cpu_unet_steps, {})
    if rank == 0 and world == 1 and not args.no_cudnn:
        del plan
        diff._plans.clear()
        torch.cuda.empty_cache()
        try:
            cud = cudnn_baseline(torch, hp, dev, min(wl["eff_batch"], 64), lr, T)
            cud["ours_over_tf32"] = value / cud["tf32"]["images_per_s"]
            cud["ours_over_bf16_channels_last"] = value / cud["bf16_channels_last"]["images_per_s"]
        except Exception as e:                      # a baseline leg must never take the bench line down
            cud = {"error": str(e).splitlines()[0][:200]}

    if rank == 0:
        h2d = int(calls * (img_lr_h.numel() + img_lr_up_h.numel()) * 4)      # summed over ranks: each uploads its shard
        d2h = int(calls * out_h.numel() * 4)                                  # rank 0 reads the gathered batch
        n_launch = calls * diff.last_stats.get("gpu_launches", 0) if not tile else diff.last_stats.get("gpu_launches", 0)
        line = {
            "metric": METRIC, "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "bf16" if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"] + ", random-init weights from srdiff/config.yaml", "name": args.workload,
                       "global_batch": G, "per_gpu_batch_per_call": micro, "sample_calls_per_step": calls,
                       "parallelism": "batch-sharded x%d via dist.sample_sharded, final NCCL all_gather" % world
                       if world > 1 else "single GPU",
                       "noise": "in-kernel Philox", "cuda_graph": True,
                       "l2": "working set per Unet pass (%.1f GB of activations) is far larger than the 126 MB L2"
                             % (MB_UNET_STEP * 1e6 * Bp * px_scale / 1e9)},
            "clocks": clocks,
            "e2e": {"value": value_e2e, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(args.steps * n_launch),
            "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (all %d tcgen05 conv layers of one Unet pass in "
                                                      "%d launches)" % (n_tc, n_ops),
                         "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"], "traffic": traffic, "traffic_unit": "bytes per Unet pass",
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_pass": MB_UNET_STEP * 1e6 * Bp * px_scale,
                         "peak_source": peaks["source"] + ", sustained bf16",
                         "flops_per_pass": flops_per_pass, "ms_per_pass": ms_unet},
            "roofline_p_sample": {"bound": "hbm", "achieved": ps_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                  "frac": ps_gbs / peaks["hbm_gbs"], "bytes_per_launch": ps_bytes,
                                  "bytes_counted": "x, eps in + x' out (fp32) + padded NHWC bf16 copy of x'",
                                  "l2": "cold: a 512 MB read evicts the kernel's 66 MB working set before every timed launch "
                                        "(graph of 20 x [flush, p_sample] minus a graph of 20 x [flush])",
                                  "us_per_launch": ms_ps * 1000.0},
            "whole_job_tensor_frac": (GFLOP_UNET_STEP * T + GFLOP_ONCE) * (H * H / 25600.0) * 1e9 * value
            / world / 1e12 / peaks["tflops"],
            "cpu_baseline": None if cpu is None else {"value": cpu["value"], "unit": "images/s",
                                                      "cores": cpu["cores"], "kind": cpu["kind"],
                                                      "sample": cpu["sample"]},
            "cudnn_baseline": cud,
        }
        if shard_inv is not None:
            line["shard_invariance"] = shard_inv
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
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="override the workload's batch (per GPU; global for c3)")
    ap.add_argument("--lr", type=int, default=None, help="override the workload's LR side")
    ap.add_argument("--timesteps", type=int, default=None, help="override the workload's T")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "bf16_direct"])
    ap.add_argument("--cpu-batch", type=int, default=64, help="images of the CPU sample (the full batch by default)")
    ap.add_argument("--cpu-unet-steps", type=int, default=1)
    ap.add_argument("--cpu-port", action="store_true", help="time the oracle port even if oracle/_ref is present")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cudnn", action="store_true")
    ap.add_argument("--layers", action="store_true", help="also dump per-layer timings to gpurun_out/layers.json")
    args = ap.parse_args()
    hp = load_hparams()
    wl = dict(WORKLOADS[args.workload])
    wl["eff_batch"] = args.batch or wl["batch"]
    if args.lr:
        wl["lr"] = args.lr
    if args.timesteps:
        wl["T"] = args.timesteps
    if args.batch or args.lr or args.timesteps:
        wl["desc"] += " [overridden: batch %d, lr %d, T %d]" % (wl["eff_batch"], wl["lr"], wl["T"])
    if args.impl == "reference":
        run_reference_arm(args, hp, wl)
    else:
        run_ours(args, hp, wl)


if __name__ == "__main__":
    main()
