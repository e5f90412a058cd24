"""GPU helper: where does a conv launch spend its time outside the steady state?  Library built with
-DSRD_TC_DEBUG=3; the same layer is launched back to back (as in the sampler) and the LAST launch is reported.
    python tools/timeline_kernel.py CIN COUT H B [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dgm_img_super_resolution_b200 import engine as E, _lib as L
cin, cout, H, B = [int(v) for v in sys.argv[1:5]]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
dev = "cuda:0"
torch.manual_seed(0)
conv = torch.nn.Conv2d(cin, cout, 3).to(dev)
dbgs = [torch.zeros((148, 64, 8), dtype=torch.int64, device=dev) for _ in range(reps)]
blds = []
for r in range(reps):
    L.lib().srd_tc_set_debug(dbgs[r].data_ptr())
    bld = E.Builder(dev, "bf16")
    src = bld.act(B, H, H, cin); src.t.normal_()
    out = bld.act(B, H, H, cout)
    bld.conv(E._pc(conv), [(src, 0, cin)], out=out, act=L.ACT_MISH, border=L.BORDER_REFLECT)
    blds.append(bld)
L.lib().srd_tc_set_debug(None)
for _ in range(3):
    for b in blds: b.run()
torch.cuda.synchronize()
for d in dbgs: d.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for b in blds: b.run()
e1.record(); torch.cuda.synchronize()
print("cin=%d cout=%d H=%d B=%d: %.1f us per launch (%d back-to-back launches)" % (cin, cout, H, B, e0.elapsed_time(e1) * 1000 / reps, reps))
D = [d.cpu().numpy()[:, 63, :] for d in dbgs]          # [cta][8]
for r in range(1, reps):
    d, p = D[r], D[r - 1]
    act = d[:, 6] > 0
    t_entry, t_exit = d[act, 6], d[act, 7]
    prev_exit = p[p[:, 7] > 0, 7]
    cyc = d[act]
    names = ["setup", "pdl_wait", "roles", "final_sync", "dealloc"]
    deltas = [np.median(cyc[:, k + 1] - cyc[:, k]) for k in range(5)]
    print("launch %d: CTAs %d | first entry %+.2f us / last entry %+.2f us rel. to prev grid's LAST exit | grid span %.2f us (first entry -> last exit)" % (
        r, act.sum(), (t_entry.min() - prev_exit.max()) / 1e3, (t_entry.max() - prev_exit.max()) / 1e3, (t_exit.max() - t_entry.min()) / 1e3))
    print("   median cycles per CTA: " + " | ".join("%s %d" % (n, v) for n, v in zip(names, deltas)) +
          " | exit spread %.2f us" % ((t_exit.max() - t_exit.min()) / 1e3))
    print("   prev last exit -> this last exit: %.2f us" % ((t_exit.max() - prev_exit.max()) / 1e3))
