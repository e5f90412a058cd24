"""GPU helper: where does the layer chain lose / win time?  Times, for the bench shape, (a) every layer as a plain
launch, (b) the same layer as a one-layer chain (no dependencies: the cost of the chain kernel's own structure),
(c) chains over runs of layers, (d) the whole chain.  Each measurement = a CUDA graph of `reps` back-to-back runs.
    python tools/chain_time.py [B] [lr] [spec ...]      spec = S:K chains ops[S:K]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgm_img_super_resolution_b200.model as M
from dgm_img_super_resolution_b200 import engine as E, _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lr = int(sys.argv[2]) if len(sys.argv) > 2 else 40
specs = sys.argv[3:]
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
plan = E.SamplerPlan(diff, B, lr, lr, "bf16", dev)
plan.img_lr.uniform_(-1, 1); plan.img_lr_up.uniform_(-1, 1); plan.x.normal_()
plan.encode(); plan.load_x(); torch.cuda.synchronize()
ops = plan.unet_prog.bld.ops
ctx = E.RunCtx(plan.temb.data_ptr())

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (2 * reps)

def name(op):
    d = op.desc
    return "%dx%d@%d s%d u%d k%d%s" % (sum(d.src[j].cin for j in range(d.n_src)), d.N, d.H_out, d.stride, d.up, d.nky * d.nkx, "+1x1" if d.N2 else "")

def run_ops(lo, hi):
    for op in ops[lo:hi]:
        op.run(ctx)

if specs and specs[0] == "one":
    for i in [int(v) for v in specs[1:]] or [2, 6, 12, 18, 37]:
        op = ops[i]
        c1 = E.ChainOp([op], fused=True)
        ta = timed(lambda: op.run(ctx), reps=5)
        tb = timed(lambda: c1.run(ctx), reps=5)
        print("%2d %-24s plain %7.1f us | 1-layer fused segment %7.1f us (%+.1f)" % (i, name(op), ta, tb, tb - ta), flush=True)
    sys.exit(0)
if specs and specs[0] == "fused":
    # per fused group of the engine's default grouping: its launches one by one vs the fused-segment launch
    bld = plan.unet_prog.bld
    E.FUSE_SEGMENTS = True
    bld.fuse_segments()
    pos = 0
    for op in bld.exec:
        if isinstance(op, E.ChainOp):
            n = len(op.ops)
            ta = timed(lambda: [o.run(ctx) for o in op.ops], reps=3)
            tb = timed(lambda: op.run(ctx), reps=3)
            one = []
            if len(specs) > 1:
                for o in op.ops[:1]:
                    c1 = E.ChainOp([o], fused=True)
                    one.append(("plain %.1f" % timed(lambda: o.run(ctx), reps=5), "fused-1 %.1f" % timed(lambda: c1.run(ctx), reps=5)))
            print("ops[%d:%d] %-22s x%d: launches %.1f us | fused %.1f us (%+.1f) dir=%s %s" % (
                pos, pos + n, name(op.ops[0]), n, ta, tb, tb - ta, [o.desc.reverse_items for o in op.ops], one), flush=True)
            pos += n
        else:
            pos += 1
    sys.exit(0)
if not specs:
    tot_a = tot_b = 0.0
    for i, op in enumerate(ops):
        ta = timed(lambda: op.run(ctx))
        ch = E.ChainOp([op])
        tb = timed(lambda: ch.run(ctx))
        tot_a += ta; tot_b += tb
        print("%2d %-24s launch %7.1f us | 1-layer chain %7.1f us (%+.1f)" % (i, name(op), ta, tb, tb - ta), flush=True)
    print("sum: launches %.1f us, 1-layer chains %.1f us" % (tot_a, tot_b))
    specs = ["1:4", "6:9", "12:15", "16:25", "18:25", "27:30", "32:35", "37:40", "0:42"]
for spec in specs:
    S, K = [int(v) for v in spec.split(":")]
    ch = E.ChainOp(ops[S:K])
    ta = timed(lambda: run_ops(S, K), reps=3)
    tb = timed(lambda: ch.run(ctx), reps=3)
    print("ops[%d:%d] (%d segments): launches %.1f us | chain %.1f us (%+.1f)" % (S, K, ch.n_segments, ta, tb, tb - ta), flush=True)
