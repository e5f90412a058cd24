"""GPU helper: per-item role timeline of a FUSED segment (library built with -DSRD_TC_DEBUG=1): every layer of the
group stamps into its own buffer; one CTA's stamps are printed on a common clock.
    python tools/timeline_fused.py S K [cta]      (layers ops[S:K] of the bench-shape Unet)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgm_img_super_resolution_b200.model as M
from dgm_img_super_resolution_b200 import engine as E, _lib as L
S, K = int(sys.argv[1]), int(sys.argv[2])
ctas = [int(v) for v in sys.argv[3:]] or [0, 140]
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
if os.environ.get("FWD") == "1":
    E.REVERSE_ORDER = False
dbgs = []
orig_init = E.ConvOp.__init__
def patched(self, desc, keep, temb_off=None, io=None):
    d = torch.zeros((148, 64, 8), dtype=torch.int64, device=dev)
    dbgs.append(d)
    L.lib().srd_tc_set_debug(d.data_ptr())
    orig_init(self, desc, keep, temb_off, io)
    L.lib().srd_tc_set_debug(None)
orig_unet = E.UnetProgram.__init__
def unet_init(self, *a, **k):
    E.ConvOp.__init__ = patched
    try:
        orig_unet(self, *a, **k)
    finally:
        E.ConvOp.__init__ = orig_init
E.UnetProgram.__init__ = unet_init
plan = E.SamplerPlan(diff, 64, 40, 40, "bf16", dev)
plan.img_lr.uniform_(-1, 1); plan.img_lr_up.uniform_(-1, 1); plan.x.normal_()
plan.encode(); plan.load_x(); torch.cuda.synchronize()
ops = plan.unet_prog.bld.ops
ctx = E.RunCtx(plan.temb.data_ptr())
plan.unet_prog.bld.run(ctx); torch.cuda.synchronize()
dirs = [o.desc.reverse_items for o in ops[S:K]]
for o in ops[S + 1:K]:                       # one direction inside the group, like Builder.fuse_segments
    if o.desc.reverse_items != ops[S].desc.reverse_items:
        o.desc.reverse_items = ops[S].desc.reverse_items
        L.lib().srd_tc_set_debug(dbgs[ops.index(o)].data_ptr()); o.rebuild(); L.lib().srd_tc_set_debug(None)
mode = os.environ.get("MODE", "fused")
ch = E.ChainOp(ops[S:K], fused=True) if mode == "fused" else None
def run():
    if ch is not None: ch.run(ctx)
    else:
        for o in ops[S:K]: o.run(ctx)
for _ in range(3): run()
torch.cuda.synchronize()
for d in dbgs: d.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("ops[%d:%d] %s: %.1f us, tile directions %s" % (S, K, mode, e0.elapsed_time(e1) * 1000, dirs))
names = ["prodA", "mma_start", "mma_Tempty", "mma_commit", "mma_Afull", "epi_Tfull", "epi_release", "epi_done"]
for cta in ctas:
    base = None
    print("CTA %d  (cycles since the first stamp)  %s" % (cta, names))
    for li in range(S, K):
        t = dbgs[li].cpu().numpy()[cta]
        n = int((t[:, 3] > 0).sum()) if (t[:, 3] > 0).any() else int((t[:, 0] > 0).sum())
        for i in range(min(n, 4)):
            vals = t[i]
            if base is None:
                base = min(int(v) for v in vals if v > 0)
            print("  layer %2d item %d: " % (li, i) + " ".join("%8d" % (int(v - base) if v > 0 else -1) for v in vals))
