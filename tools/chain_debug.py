"""GPU helper: run the first K layers of the Unet launch list as a chain (the rest layer by layer) and compare with
the all-layer-by-layer result; on a failure print the pipeline watchdog code.
    python tools/chain_debug.py B h w [K ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgm_img_super_resolution_b200.model as M
from dgm_img_super_resolution_b200 import engine as E, _lib as L

B, h, w = [int(v) for v in sys.argv[1:4]]
Ks = [v for v in sys.argv[4:]] or ["0:2"]          # "S:K" chain over ops[S:K]; "S:K:r" = run it twice, no re-zeroing
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hp = json.load(open(os.path.join(ROOT, "tests", "golden", "hparams.json")))
hp["unet_dim_mults"] = tuple(hp["unet_dim_mults"])
M.hparams.clear(); M.hparams.update(hp)
dev = torch.device("cuda:0")
torch.manual_seed(1234)
nf = hp["rrdb_num_feat"]
rrdb = M.RRDBNet(3, 3, nf, hp["rrdb_num_block"], nf // 2)
unet = M.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
diff = M.GaussianDiffusion(unet, rrdb, timesteps=4).to(dev).eval()
E.USE_CHAIN = False
plan = E.SamplerPlan(diff, B, h, w, "bf16", dev)
plan.img_lr.uniform_(-1, 1); plan.img_lr_up.uniform_(-1, 1); plan.x.normal_()
plan.encode(); plan.load_x(); torch.cuda.synchronize()
bld = plan.unet_prog.bld
ops = bld.ops
ctx = E.RunCtx(plan.temb.data_ptr())
for i, op in enumerate(ops):
    d = op.desc
    print(i, "cin", sum(d.src[j].cin for j in range(d.n_src)), "N", d.N, "H", d.H_out, "s", d.stride, "up", d.up, "k", d.nky * d.nkx,
          "N2", d.N2, "res1", bool(d.res1), "res2", bool(d.res2))
bld.run(ctx); torch.cuda.synchronize()
ref = plan.unet_prog.eps.clone()
acts = {}
def snapshot():
    return [op.io["out"].t.clone() if op.io["out"] is not None else None for op in ops]
ref_acts = snapshot()
print("chainable:", E.ChainOp.chainable(ops))
for spec in Ks:
    repeat = spec.endswith(":r")
    parts = [int(v) for v in spec.replace(":r", "").split(":")]
    S, K = (0, parts[0]) if len(parts) == 1 else parts[:2]
    for op in ops:
        if op.io["out"] is not None: op.io["out"].t.zero_()
        if op.io["dst2"] is not None: op.io["dst2"].t.zero_()
    plan.unet_prog.eps.fill_(float("nan"))
    try:
        ch = E.ChainOp(ops[S:K])
        print("chain [%d:%d] segments=%d" % (S, K, ch.n_segments), flush=True)
        for _ in range(2 if repeat else 1):
            for op in ops[:S]:
                op.run(ctx)
            ch.run(ctx)
            torch.cuda.synchronize()
            for op in ops[K:]:
                op.run(ctx)
            torch.cuda.synchronize()
    except Exception as e:
        print("K=%d FAILED: %s" % (K, str(e).splitlines()[0]))
        print("watchdog flag:", L.lib().srd_watchdog_flag())
        sys.exit(1)
    got = snapshot()
    bad = [i for i in range(len(ops)) if ref_acts[i] is not None and not torch.equal(ref_acts[i], got[i])]
    print("[%d:%d] eps equal: %s, first differing layer outputs: %s" % (S, K, torch.equal(ref, plan.unet_prog.eps), bad[:8]), flush=True)
    if bad:
        i = bad[0]
        d = (ref_acts[i].float() - got[i].float()).abs()          # [B][H+2][W+2][C]
        rows = d.amax(dim=(2, 3))
        print("   layer %d: max |diff| %.4g; per (image, padded row) max diff:" % (i, float(d.max())))
        for b in range(rows.shape[0]):
            print("   b=%d " % b + " ".join("%.2g" % float(v) for v in rows[b]))
        cols = d.amax(dim=(0, 1, 3))
        print("   per padded col: " + " ".join("%.2g" % float(v) for v in cols))
        ch_ = d.amax(dim=(0, 1, 2))
        print("   per channel (first 16): " + " ".join("%.2g" % float(v) for v in ch_[:16]))
