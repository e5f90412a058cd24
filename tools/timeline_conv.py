"""GPU helper: run one conv with the debug timeline enabled and print per-item stage timings (cycles).
    python tools/timeline_conv.py CIN COUT H B [res|convT|s2]   (H = input size)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgm_img_super_resolution_b200 import engine as E, _lib as L
cin, cout, H, B = [int(v) for v in sys.argv[1:5]]
kind = sys.argv[5] if len(sys.argv) > 5 else ""
use_res = kind == "res"
dev = "cuda:0"
torch.manual_seed(0)
if kind == "convT":
    conv, Ho, act = torch.nn.ConvTranspose2d(cin, cout, 4, 2, 1).to(dev), 2 * H, L.ACT_NONE
elif kind == "s2":
    conv, Ho, act = torch.nn.Conv2d(cin, cout, 3, 2).to(dev), H // 2, L.ACT_NONE
else:
    conv, Ho, act = torch.nn.Conv2d(cin, cout, 3).to(dev), H, L.ACT_MISH
dbg = torch.zeros((148, 64, 8), dtype=torch.int64, device=dev)
L.lib().srd_tc_set_debug(dbg.data_ptr())
bld = E.Builder(dev, "bf16")
src = bld.act(B, H, H, cin); src.t.normal_()
out = bld.act(B, Ho, Ho, cout); res = bld.act(B, Ho, Ho, cout)
temb = torch.randn(cout, device=dev)
op = bld.conv(E._pc(conv), [(src, 0, cin)], out=out, act=act, res1=(res, 0, 1.0) if use_res else None, border=L.BORDER_REFLECT)
L.lib().srd_tc_set_debug(None)
for _ in range(3):
    bld.run()
torch.cuda.synchronize()
dbg.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); bld.run(); e1.record(); torch.cuda.synchronize()
print("kernel %.1f us" % (e0.elapsed_time(e1) * 1000))
d = dbg.cpu().numpy()
names = ["prodA", "mma_start", "mma_Tempty", "mma_commit", "mma_Afull", "epi_Tfull", "epi_release", "epi_done"]
for cta in (0, 73, 147):
    t = d[cta]
    n = int((t[:, 3] > 0).sum())
    base = t[0, 0] if t[0, 0] > 0 else t[0, 1]
    print("CTA %d: %d items; stamps relative to first (cycles): %s" % (cta, n, names))
    for i in range(min(n, 14)):
        print("  item %2d: " % i + " ".join("%7d" % (int(v - base) if v > 0 else -1) for v in t[i]))
    if n > 4:
        import numpy as np
        per = (t[n - 1, 3] - t[1, 3]) / (n - 2)
        print("  steady-state period (mma_commit to mma_commit): %.0f cycles/item" % per)
        print("  mean mma issue (Tempty->commit): %.0f ; mean epi (Tfull->done): %.0f ; Tfull - commit: %.0f ; Afull wait (start->Afull): %.0f" % (
            np.mean(t[1:n, 3] - t[1:n, 2]), np.mean(t[1:n, 7] - t[1:n, 5]), np.mean(t[1:n, 5] - t[1:n, 3]), np.mean(t[1:n, 4] - t[1:n, 2])))
