"""GPU helper: per-item timeline of ONE epilogue warp (library built with -DSRD_TC_DEBUG=2).
    python tools/timeline_epi.py CIN COUT H B [conv3|convT2] [res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dgm_img_super_resolution_b200 import engine as E, _lib as L
cin, cout, H, B = [int(v) for v in sys.argv[1:5]]
kind = sys.argv[5] if len(sys.argv) > 5 else "conv3"
use_res = "res" in sys.argv[6:]
dev = "cuda:0"
torch.manual_seed(0)
conv = (torch.nn.ConvTranspose2d(cin, cout, 4, 2, 1) if kind == "convT2" else torch.nn.Conv2d(cin, cout, 3)).to(dev)
dbg = torch.zeros((148, 64, 8), dtype=torch.int64, device=dev)
L.lib().srd_tc_set_debug(dbg.data_ptr())
bld = E.Builder(dev, "bf16")
src = bld.act(B, H, H, cin, slack=64); src.t.normal_()
Ho = 2 * H if kind == "convT2" else H
out = bld.act(B, Ho, Ho, cout); res = bld.act(B, Ho, Ho, cout)
op = bld.conv(E._pc(conv), [(src, 0, cin)], out=out, act=L.ACT_NONE if kind == "convT2" else L.ACT_MISH,
              res1=(res, 0, 1.0) if use_res else None, border=L.BORDER_REFLECT, win8=(cin <= 4))
L.lib().srd_tc_set_debug(None)
for _ in range(3):
    bld.run()
torch.cuda.synchronize()
dbg.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); bld.run(); e1.record(); torch.cuda.synchronize()
print("%s cin=%d cout=%d H=%d B=%d res=%s: kernel %.1f us" % (kind, cin, cout, H, B, use_res, e0.elapsed_time(e1) * 1000))
d = dbg.cpu().numpy()
names = ["top", "Tfull", "buf", "chunks", "fence", "store", "drain", "prep"]
for cta in (0, 2):
    t = d[cta]
    idx = [i for i in range(64) if t[i, 1] > 0]
    print("CTA %d: %d stamped items; deltas (cycles) between consecutive stamps: %s" % (cta, len(idx), names))
    rows = []
    for i in idx[:40]:
        r = t[i]
        rows.append([int(r[k + 1] - r[k]) for k in range(7)])
    for i, r in zip(idx[:12], rows[:12]):
        print("  item %2d: wait_Tfull %6d | wait_buf %6d | chunks %6d | fence %5d | store %5d | drain %6d | prep %5d" % (i, *r))
    if rows:
        m = np.array(rows[2:]).mean(0)
        print("  mean    : wait_Tfull %6.0f | wait_buf %6.0f | chunks %6.0f | fence %5.0f | store %5.0f | drain %6.0f | prep %5.0f  (sum %.0f)" % (*m, m.sum()))
        same = [a_ for a_ in idx if (a_ - idx[0]) % 4 == 0]
        if len(same) > 3:
            print("  same-group item spacing (top to top): %.0f cycles" % np.mean([t[same[j + 1], 0] - t[same[j], 0] for j in range(1, len(same) - 1)]))
