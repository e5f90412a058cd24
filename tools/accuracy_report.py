"""GPU helper: print the parity margins of the CUDA path against the committed reference fixtures
(tests/golden/*.pt, produced by oracle/make_golden.py from the unmodified reference):
Unet eps rel-L2 at three shapes, RRDB rel-L2, final-image PSNR of a T=4 sample().
    python tools/accuracy_report.py [bf16|fp32|bf16_direct ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from conftest import build_ours, golden, load_hparams, gen_inputs  # noqa: E402
from oracle import srdiff_oracle as O  # noqa: E402
from oracle.ref_loader import inject_noise  # noqa: E402

DEV = "cuda:0"
hp = load_hparams()
precs = sys.argv[1:] or ["bf16"]
for prec in precs:
    M, diff = build_ours(hp, device=DEV)
    M.runtime.precision = prec
    for name in ("nets_small.pt", "nets_odd.pt"):
        g = golden(name)
        feas = [f.to(DEV) for f in g["feas"]]
        eps = diff.denoise_fn(g["x"].to(DEV), g["t"].to(DEV), feas, g["img_lr_up"].to(DEV))
        out, f2 = diff.rrdb(g["img_lr"].to(DEV), True)
        print("%-11s %-14s unet eps rel-L2 %.3e | rrdb feas rel-L2 %.3e out %.3e" % (
            prec, name, O.rel_l2(eps.cpu(), g["eps"]), O.rel_l2(torch.stack(f2).cpu(), g["feas"]),
            O.rel_l2(out.cpu(), g["rrdb_out"])))
    g = golden("nets_c1.pt")
    img_lr, img_lr_up, gen = gen_inputs(1, 40, 40, hp["sr_scale"], g["seed_inputs"])
    x = torch.randn((1, 3, 160, 160), generator=gen)
    out, feas = diff.rrdb(img_lr.to(DEV), True)
    eps = diff.denoise_fn(x.to(DEV), g["t"].to(DEV), feas, img_lr_up.to(DEV))
    print("%-11s %-14s unet eps rel-L2 %.3e (configs[0] shape, fp16 fixture)" % (prec, "nets_c1.pt", O.rel_l2(eps.cpu(), g["eps"].float())))
    M4, d4 = build_ours(hp, timesteps=4, device=DEV)
    M4.runtime.precision, M4.runtime.graph, M4.runtime.noise = prec, True, "torch"
    g = golden("sample_T4.pt")
    with inject_noise(M4, g["noises"]):
        img, rrdb_out = d4.sample(g["img_lr"].to(DEV), g["img_lr_up"].to(DEV), tuple(g["img_lr_up"].shape))
    print("%-11s %-14s sample() PSNR vs reference %.1f dB, rrdb_out rel-L2 %.3e" % (
        prec, "sample_T4.pt", O.psnr(img.cpu(), g["img"]), O.rel_l2(rrdb_out.cpu(), g["rrdb_out"])))
    M4.runtime.precision = "bf16"
    # the full configs[0] trajectory: all 100 steps against the reference's final image
    from oracle.make_golden_T100 import make_noises  # noqa: E402
    g = golden("sample_T100.pt")
    M100, d100 = build_ours(hp, device=DEV)
    M100.runtime.precision, M100.runtime.graph, M100.runtime.noise = prec, True, "torch"
    img_lr, img_lr_up, _ = gen_inputs(1, 40, 40, hp["sr_scale"], g["seed_inputs"])
    with inject_noise(M100, make_noises(g["T"], img_lr_up.shape)):
        img, _ = d100.sample(img_lr.to(DEV), img_lr_up.to(DEV), tuple(img_lr_up.shape))
    print("%-11s %-14s sample() T=100 PSNR vs reference %.1f dB" % (
        prec, "sample_T100.pt", O.psnr(img.cpu(), g["img"].float())))
    M100.runtime.precision = "bf16"
