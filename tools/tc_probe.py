"""GPU diagnostic (not a test): run single convolutions on the tcgen05 kernel and on the direct
kernel against a torch fp32 reference and print error statistics per configuration.  Each
configuration runs in its own process so a faulting kernel does not poison the others.

    python tools/tc_probe.py            # all configurations
    python tools/tc_probe.py 3          # one configuration in-process
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = [
    # name, kind, cins, cout, H, W, B
    ("c64_s1", "conv3", [64], 64, 16, 16, 2),
    ("c64_s1_big", "conv3", [64], 64, 160, 160, 2),
    ("c3_s1", "conv3", [3], 64, 16, 24, 1),
    ("c128_192", "conv3", [128], 192, 20, 20, 3),
    ("c256_256", "conv3", [256], 256, 20, 20, 4),
    ("cat_256_64", "conv3", [128, 128], 64, 40, 40, 2),
    ("c1x1", "conv1", [64], 128, 16, 16, 2),
    ("s2_64", "conv3s2", [64], 64, 16, 16, 2),
    ("s2_192", "conv3s2", [192], 192, 40, 40, 2),
    ("up2_64", "convT2", [64], 64, 8, 8, 2),
    ("up2_192", "convT2", [192], 192, 20, 20, 2),
    ("up4_cond", "convT4", [32, 32, 32], 64, 10, 10, 2),
    ("rrdb_48_16", "conv3z", [48], 16, 12, 12, 2),
    ("rrdb_96_32", "conv3z", [96], 32, 40, 40, 2),
    ("n128_mt2", "conv3", [128], 128, 80, 80, 8),
]


def run_one(idx):
    import torch
    import torch.nn.functional as F
    from dgm_img_super_resolution_b200 import engine as E, _lib as L
    name, kind, cins, cout, H, W, B = CONFIGS[idx]
    dev = "cuda:0"
    torch.manual_seed(idx)
    I = sum(cins)
    x = torch.randn(B, I, H, W)
    reflect = kind in ("conv3", "conv3s2", "conv1")
    if kind.startswith("convT"):
        S = int(kind[-1])
        conv = torch.nn.ConvTranspose2d(I, cout, 2 * S, S, S // 2)
        want = conv(x)
    elif kind == "conv1":
        conv = torch.nn.Conv2d(I, cout, 1)
        want = conv(x)
    else:
        stride = 2 if kind == "conv3s2" else 1
        conv = torch.nn.Conv2d(I, cout, 3, stride)
        want = conv(F.pad(x, (1, 1, 1, 1), mode="reflect" if reflect else "constant"))
    want = want.detach()
    conv = conv.to(dev)
    res = {}
    for label, prec, kernel in (("direct", "bf16_direct", None), ("tc", "bf16", L.KERNEL_TC)):
        bld = E.Builder(dev, prec)
        srcs, a = [], 0
        for c in cins:
            act = bld.act(B, H, W, c)
            xs = x[:, a:a + c].contiguous().to(dev)
            L.check(L.lib().srd_nchw_to_act(xs.data_ptr(), B, c, H, W, 1.0, 0.0, act.ptr, act.ct, bld.prec.dtype,
                                            L.BORDER_REFLECT if reflect else L.BORDER_KEEP, E._stream()))
            srcs.append((act, 0, c))
            a += c
        out = bld.act(B, want.shape[2], want.shape[3], cout)
        op = bld.conv(E._pc(conv), srcs, out=out, kernel=kernel)
        o = torch.empty(want.shape, device=dev)
        bld.act_to_nchw(out, 0, cout, o)
        bld.run()
        torch.cuda.synchronize()
        got = o.cpu()
        err = (got - want).double()
        rel = float(err.norm() / want.double().norm())
        res[label] = rel
        bad = (err.abs() > 0.05 * want.abs().max()).float()
        print("  %-7s kernel=%d rel_l2=%.3e max_abs=%.3e bad_frac=%.4f nan=%d" %
              (label, op.kernel, rel, float(err.abs().max()), float(bad.mean()), int(torch.isnan(got).sum())))
        if rel > 2e-2:
            # where is it wrong?  per-channel / per-row / per-col error energy
            e2 = err ** 2
            print("    err by out-channel (first 16):", [round(float(v), 3) for v in e2.sum((0, 2, 3))[:16]])
            print("    err by row:", [round(float(v), 2) for v in e2.sum((0, 1, 3))[:24]])
            print("    err by col:", [round(float(v), 2) for v in e2.sum((0, 1, 2))[:24]])
            print("    err by image:", [round(float(v), 2) for v in e2.sum((1, 2, 3))])
    print("  watchdog:", L.lib().srd_watchdog_flag())
    ok = res["tc"] < 1e-2 and res["direct"] < 1e-2
    print("RESULT %-12s %s" % (name, "OK" if ok else "FAIL"))
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(run_one(int(sys.argv[1])))
    fails = 0
    for i, cfg in enumerate(CONFIGS):
        print("== %d %s" % (i, cfg), flush=True)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True, text=True,
                               timeout=180)
            print(p.stdout[-3000:], end="")
            if p.returncode != 0:
                fails += 1
                print("  rc=%d stderr tail: %s" % (p.returncode, p.stderr[-800:]))
        except subprocess.TimeoutExpired:
            fails += 1
            print("  TIMEOUT")
    print("probe: %d/%d configurations failed" % (fails, len(CONFIGS)))
