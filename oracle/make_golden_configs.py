"""TEST INFRASTRUCTURE ONLY -- fixtures at the FULL shapes of BASELINE.json configs[1] / configs[3] (batch 64 of
40x40 -> 160x160: the bench shape) and configs[4] (64x64 -> 256x256 patches), produced by running the UNMODIFIED
reference (`/root/reference/srdiff/model.py`) on the CPU of the build container.

    python -m oracle.make_golden_configs      ->  tests/golden/configs_c1.pt, tests/golden/configs_c4.pt

Per configuration: the denoiser output eps of one `Unet.forward` call at several timesteps and the final image of a
short `sample()` run with injected noise.  Only a few images of the batch are stored in full (fp16; first, middle,
last -- the last one sits at the tail of the stacked-row tiling); for ALL images of the batch the per-image L2 norm
and sum are stored, so a layout / batching bug anywhere in the batch is caught.  Inputs, weights and noise are never
stored: the tests regenerate them from the same seeds (`gen_inputs`, `torch.manual_seed(1234)` + constructors).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import OUT, SEED, gen_inputs  # noqa: E402

CASES = {
    # name: (B, lr, seed_inputs, timesteps of the single denoiser calls, T of the short sampler run, kept images)
    "c1": dict(B=64, lr=40, seed=101, ts=(99, 50, 0), T_sample=3, keep=(0, 31, 63)),
    "c4": dict(B=8, lr=64, seed=103, ts=(50,), T_sample=2, keep=(0, 7)),
}


def case_inputs(c, scale):
    """(img_lr, img_lr_up, x for the denoiser calls, noises [T_sample+1, ...] for the sampler run)."""
    img_lr, img_lr_up, g = gen_inputs(c["B"], c["lr"], c["lr"], scale, c["seed"])
    x = torch.randn(tuple(img_lr_up.shape), generator=g)
    noises = torch.randn((c["T_sample"] + 1,) + tuple(img_lr_up.shape), generator=g)
    return img_lr, img_lr_up, x, noises


def per_image(t):
    f = t.double().flatten(1)
    return {"norm": f.norm(dim=1).float(), "sum": f.sum(dim=1).float()}


def main():
    hp = ref_loader.load_hparams()
    for name, c in CASES.items():
        ref, diff = ref_loader.build_reference(hp, seed=SEED)          # T = 100 (config.yaml)
        img_lr, img_lr_up, x, noises = case_inputs(c, hp["sr_scale"])
        keep = list(c["keep"])
        out = {"case": dict(c), "eps": {}, "eps_stats": {}}
        with torch.no_grad():
            rrdb_out, feas = diff.rrdb(img_lr, True)
            out["rrdb_out"] = rrdb_out[keep].half()
            out["rrdb_out_stats"] = per_image(rrdb_out)
            for t in c["ts"]:
                tt = torch.full((c["B"],), t, dtype=torch.long)
                eps = diff.denoise_fn(x, tt, feas, img_lr_up)
                out["eps"][t] = eps[keep].half()
                out["eps_stats"][t] = per_image(eps)
                print(name, "t=%d" % t, "|eps| %.4f" % float(eps.norm()), flush=True)
        ref, diffS = ref_loader.build_reference(hp, timesteps=c["T_sample"], seed=SEED)
        with ref_loader.inject_noise(ref, noises):
            with torch.no_grad():
                img, _ = diffS.sample(img_lr, img_lr_up, tuple(img_lr_up.shape))
        out["img"] = img[keep].half()
        out["img_stats"] = per_image(img)
        p = os.path.join(OUT, "configs_%s.pt" % name)
        torch.save(out, p)
        print("%s %d bytes" % (p, os.path.getsize(p)), flush=True)


if __name__ == "__main__":
    main()
