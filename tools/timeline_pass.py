"""GPU helper: per-launch phases of ONE WHOLE Unet pass in its real sequence, under sustained load.
Library built with -DSRD_TC_DEBUG=3 (tools/build_variants.sh dbg3="-DSRD_TC_DEBUG=3", SRD_LIB_PATH=...).
Every conv plan gets its own stamp buffer; the pass is replayed `reps` times back to back (so the chip sits at
its sustained clocks) and the LAST pass is reported: for each launch the period (previous launch's last CTA exit
-> this launch's last exit), the median role time of a CTA, and how the rest splits up.
    python tools/timeline_pass.py [B] [lr] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import dgm_img_super_resolution_b200.model as M
from dgm_img_super_resolution_b200 import engine as E, _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lr = int(sys.argv[2]) if len(sys.argv) > 2 else 40
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
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

dbgs = []
orig_init = E.ConvOp.__init__
def patched(self, desc, keep, temb_off=None):
    d = torch.zeros((148, 64, 8), dtype=torch.int64, device=dev)
    dbgs.append(d)
    L.lib().srd_tc_set_debug(d.data_ptr())
    orig_init(self, desc, keep, temb_off)
    L.lib().srd_tc_set_debug(None)
prog_holder = {}
orig_unet = E.UnetProgram.__init__
def unet_init(self, *a, **k):
    E.ConvOp.__init__ = patched
    try:
        orig_unet(self, *a, **k)
    finally:
        E.ConvOp.__init__ = orig_init
E.UnetProgram.__init__ = unet_init
plan = E.SamplerPlan(diff, B, lr, lr, "bf16", dev)
plan.img_lr.uniform_(-1, 1); plan.img_lr_up.uniform_(-1, 1); plan.x.normal_()
plan.encode(); plan.load_x()
torch.cuda.synchronize()
temb = plan.temb.data_ptr()
g = torch.cuda.CUDAGraph()
plan.unet_prog.run(temb); torch.cuda.synchronize()
with torch.cuda.graph(g):
    plan.unet_prog.run(temb)
for _ in range(reps):
    g.replay()
torch.cuda.synchronize()
for d in dbgs: d.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(reps):
    g.replay()
e0.record(); g.replay(); e1.record()
torch.cuda.synchronize()
print("Unet pass (graph, sustained): %.1f us" % (e0.elapsed_time(e1) * 1000))
ops = [op for op in plan.unet_prog.bld.ops if isinstance(op, E.ConvOp)]
D = [d.cpu().numpy()[:, 63, :] for d in dbgs]
prev_exit = None
tot_period = tot_roles = 0.0
print("%3s %-22s %8s %8s %8s | %7s %7s %7s %7s | %7s %7s %6s" % (
    "#", "layer", "period", "roles", "ovh", "setup", "pdlwait", "fsync", "spread", "entry-", "GHz", "grid"))
rows = []
for i, (op, d) in enumerate(zip(ops, D)):
    act = d[:, 6] > 0
    if not act.any():
        continue
    c = d[act]
    t_entry, t_exit = c[:, 6], c[:, 7]
    ghz = np.median((c[:, 5] - c[:, 0]) / np.maximum(1, (t_exit - t_entry)))
    roles_us = np.median(c[:, 3] - c[:, 2]) / ghz / 1e3
    roles_max_us = np.max(c[:, 3] - c[:, 2]) / ghz / 1e3
    setup_us = np.median(c[:, 1] - c[:, 0]) / ghz / 1e3
    pdl_us = np.median(c[:, 2] - c[:, 1]) / ghz / 1e3
    fs_us = np.median(c[:, 4] - c[:, 3]) / ghz / 1e3
    spread = (t_exit.max() - t_exit.min()) / 1e3
    dd = op.desc
    name = "%dx%d@%d s%d u%d k%d" % (sum(dd.src[j].cin for j in range(dd.n_src)), dd.N, dd.H_out, dd.stride, dd.up, dd.nky * dd.nkx)
    if prev_exit is not None:
        period = (t_exit.max() - prev_exit) / 1e3
        early = (prev_exit - t_entry.min()) / 1e3
        tot_period += period; tot_roles += roles_us
        print("%3d %-22s %8.1f %8.1f %8.1f | %7.2f %7.2f %7.2f %7.2f | %7.2f %7.3f %6d" % (
            i, name, period, roles_us, period - roles_us, setup_us, pdl_us, fs_us, spread, early, ghz, act.sum()))
        rows.append(dict(i=i, name=name, period=period, roles=roles_us, roles_max=roles_max_us, setup=setup_us,
                         pdl=pdl_us, fsync=fs_us, spread=spread, early=early, ghz=float(ghz)))
    prev_exit = t_exit.max()
print("sum of periods %.1f us, sum of median role times %.1f us, overhead %.1f us (%.1f %%)" % (
    tot_period, tot_roles, tot_period - tot_roles, 100 * (tot_period - tot_roles) / tot_period))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "timeline_pass.json"), "w"), indent=1)
