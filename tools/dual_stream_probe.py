"""GPU experiment: does interleaving two independent half-batches on two streams fill the per-launch tail / fill
bubbles of the persistent conv kernels?  Compares one B-image graph against two B/2-image graphs replayed
concurrently on two streams.   python tools/dual_stream_probe.py [B] [splits]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import build_ours, load_hparams, gen_inputs
from dgm_img_super_resolution_b200 import engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
splits = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda:0"
hp = load_hparams()
M, diff = build_ours(hp, device=dev)
M.runtime.noise, M.runtime.precision = "philox", "bf16"


def make(b, off):
    plan = E.SamplerPlan(diff, b, 40, 40, "bf16", torch.device(dev))
    img_lr, img_lr_up, _ = gen_inputs(b, 40, 40, 4, 5 + off)
    plan.img_lr.copy_(img_lr.to(dev)); plan.img_lr_up.copy_(img_lr_up.to(dev))
    plan.x.normal_()
    return plan


def timed(plans, streams, reps=3):
    for p, s in zip(plans, streams):                      # capture
        with torch.cuda.stream(s):
            p.run_graph(False, 1234, 0)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for p, s in zip(plans, streams):
            with torch.cuda.stream(s):
                p.run_graph(False, 1234, 0)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


one = timed([make(B, 0)], [torch.cuda.Stream()])
print("1 x %d images: %.1f ms -> %.1f images/s" % (B, one, B / one * 1e3))
plans = [make(B // splits, i) for i in range(splits)]
two = timed(plans, [torch.cuda.Stream() for _ in range(splits)])
print("%d x %d images on %d streams: %.1f ms -> %.1f images/s (%.1f %%)" % (splits, B // splits, splits, two, B / two * 1e3, 100 * (one / two - 1)))
