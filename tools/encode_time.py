"""GPU helper: time of the once-per-image part (RRDB encoder + cond_proj [+ folded Downsample]) for a batch.
    python tools/encode_time.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import build_ours, load_hparams, gen_inputs
from dgm_img_super_resolution_b200 import engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda:0"
hp = load_hparams()
M, diff = build_ours(hp, device=dev)
plan = E.SamplerPlan(diff, B, 40, 40, "bf16", torch.device(dev))
img_lr, img_lr_up, _ = gen_inputs(B, 40, 40, 4, 5)
plan.img_lr.copy_(img_lr.to(dev)); plan.img_lr_up.copy_(img_lr_up.to(dev))
for _ in range(3):
    plan.encode()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    plan.encode()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g.replay(); torch.cuda.synchronize()
e0.record()
for _ in range(5):
    g.replay()
e1.record(); torch.cuda.synchronize()
n_r, n_tc = plan.rrdb_prog.bld.count_kernels()
print("encode (graph replay): %.3f ms per batch of %d; RRDB launches %d (%d on the tcgen05 kernel), cond launches %d" % (
    e0.elapsed_time(e1) / 5, B, n_r, n_tc, plan.cond_prog.bld.count_kernels()[0]))
