"""GPU helper for ncu: run layer L of the bench-shape Unet either as a plain launch or as a one-layer chain.
    python tools/profile_chain.py L chain|plain [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgm_img_super_resolution_b200.model as M
from dgm_img_super_resolution_b200 import engine as E, _lib as L
Li, mode = int(sys.argv[1]), sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hp = json.load(open(os.path.join(ROOT, "tests", "golden", "hparams.json")))
hp["unet_dim_mults"] = tuple(hp["unet_dim_mults"])
M.hparams.clear(); M.hparams.update(hp)
dev = torch.device("cuda:0")
torch.manual_seed(1234)
nf = hp["rrdb_num_feat"]
rrdb = M.RRDBNet(3, 3, nf, hp["rrdb_num_block"], nf // 2)
unet = M.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
diff = M.GaussianDiffusion(unet, rrdb, timesteps=100).to(dev).eval()
E.USE_CHAIN = False
plan = E.SamplerPlan(diff, 64, 40, 40, "bf16", dev)
plan.img_lr.uniform_(-1, 1); plan.img_lr_up.uniform_(-1, 1); plan.x.normal_()
plan.encode(); plan.load_x()
ops = plan.unet_prog.bld.ops
ctx = E.RunCtx(plan.temb.data_ptr())
plan.unet_prog.bld.run(ctx); torch.cuda.synchronize()
if mode == "fused":
    ch = E.ChainOp(ops[Li:Li + 1], fused=True)
    torch.cuda.profiler.start()
    for _ in range(reps):
        ch.run(ctx)
elif mode == "chain":
    torch.cuda.profiler.start()
    lo, hi = (Li, Li + 1) if len(sys.argv) <= 4 else (Li, int(sys.argv[4]))
    ch = E.ChainOp(ops[lo:hi])
    for _ in range(reps):
        ch.run(ctx)
else:
    torch.cuda.profiler.start()
    for _ in range(reps):
        ops[Li].run(ctx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
