"""GPU helper (not a test): run ONE fused conv layer a few times so ncu can capture it.
    python tools/profile_conv.py CIN COUT H B [kind=conv3|conv1|conv3s2|convT2] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dgm_img_super_resolution_b200 import engine as E, _lib as L  # noqa: E402

cin, cout, H, B = [int(v) for v in sys.argv[1:5]]
kind = sys.argv[5] if len(sys.argv) > 5 else "conv3"
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = "cuda:0"
torch.manual_seed(0)
if kind == "convT2":
    conv = torch.nn.ConvTranspose2d(cin, cout, 4, 2, 1).to(dev)
elif kind == "conv1":
    conv = torch.nn.Conv2d(cin, cout, 1).to(dev)
else:
    conv = torch.nn.Conv2d(cin, cout, 3, 2 if kind == "conv3s2" else 1).to(dev)
bld = E.Builder(dev, "bf16")
src = bld.act(B, H, H, cin, slack=64)
src.t.normal_()
Ho = H * 2 if kind == "convT2" else (H // 2 if kind == "conv3s2" else H)
out = bld.act(B, Ho, Ho, cout)
res = bld.act(B, Ho, Ho, cout)
temb = torch.randn(cout, device=dev)
nores = os.environ.get("NORES") == "1"
op = bld.conv(E._pc(conv), [(src, 0, cin)], out=out, act=L.ACT_MISH, res1=None if (nores or cin <= 4) else (res, 0, 1.0),
              border=L.BORDER_REFLECT, win8=(cin <= 4 and kind == "conv3"))
op.desc.temb = temb.data_ptr()
for _ in range(2):
    bld.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    bld.run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / reps
taps = {"conv3": 9, "conv3s2": 9, "conv1": 1, "convT2": 4}[kind]
fl = 2.0 * B * Ho * Ho * cout * cin * taps
print("kind=%s cin=%d cout=%d H=%d B=%d kernel=%d: %.1f us, %.1f TFLOP/s" % (kind, cin, cout, H, B, op.kernel, us, fl / us / 1e6))
