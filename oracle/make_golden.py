"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference/srdiff/model.py, imported through oracle/ref_loader.py) in the build container.

    python -m oracle.make_golden

The fixtures are what pins both the oracle restatement and the CUDA path on the GPU box, where
/root/reference does not exist.  Weights are never stored: every consumer rebuilds them with
`torch.manual_seed(1234)` + the constructors (the construction order consumes the CPU RNG
identically in the reference and in dgm_img_super_resolution_b200.model -- checked by
tests/test_api.py and by the per-key checksums stored here).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 1234


def gen_inputs(B, h, w, scale, seed):
    g = torch.Generator().manual_seed(seed)
    img_lr = torch.rand((B, 3, h, w), generator=g) * 2 - 1
    img_lr_up = torch.nn.functional.interpolate(img_lr, scale_factor=scale, mode="bicubic", align_corners=True)
    return img_lr, img_lr_up, g


def main():
    os.makedirs(OUT, exist_ok=True)
    hp = ref_loader.load_hparams()
    with open(os.path.join(OUT, "hparams.json"), "w") as f:
        json.dump(hp, f, indent=1, sort_keys=True)

    # ---- schedules (row G) ---------------------------------------------------------------
    sched = {}
    for name, T, over in (("cosine_T100", 100, {}), ("cosine_T25", 25, {}), ("cosine_T1000", 1000, {}),
                          ("linear_T100", 100, {"beta_schedule": "linear"})):
        h2 = dict(hp, **over)
        ref, diff = ref_loader.build_reference(h2, timesteps=T, seed=SEED)
        sched[name] = {k: v.clone() for k, v in diff.state_dict().items() if "." not in k}
    torch.save(sched, os.path.join(OUT, "schedules.pt"))

    # ---- weights checksum ----------------------------------------------------------------
    ref, diff = ref_loader.build_reference(hp, seed=SEED)
    sd = diff.state_dict()
    chk = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}
    with open(os.path.join(OUT, "state_checksum.json"), "w") as f:
        json.dump(chk, f)

    # ---- RRDB + Unet on a tiny batch (B=2, 8x8 -> 32x32) ----------------------------------
    img_lr, img_lr_up, g = gen_inputs(2, 8, 8, hp["sr_scale"], 7)
    x = torch.randn((2, 3, 32, 32), generator=g)
    t = torch.tensor([3, 57], dtype=torch.long)
    with torch.no_grad():
        rrdb_out, feas = diff.rrdb(img_lr, True)
        eps = diff.denoise_fn(x, t, feas, img_lr_up)
    torch.save({"img_lr": img_lr, "img_lr_up": img_lr_up, "x": x, "t": t, "rrdb_out": rrdb_out,
                "feas": torch.stack(feas), "eps": eps}, os.path.join(OUT, "nets_small.pt"))

    # ---- odd shape (B=3, 6x10 -> 24x40) ---------------------------------------------------
    img_lr, img_lr_up, g = gen_inputs(3, 6, 10, hp["sr_scale"], 11)
    x = torch.randn((3, 3, 24, 40), generator=g)
    t = torch.tensor([0, 99, 42], dtype=torch.long)
    with torch.no_grad():
        rrdb_out, feas = diff.rrdb(img_lr, True)
        eps = diff.denoise_fn(x, t, feas, img_lr_up)
    torch.save({"img_lr": img_lr, "img_lr_up": img_lr_up, "x": x, "t": t, "rrdb_out": rrdb_out,
                "feas": torch.stack(feas), "eps": eps}, os.path.join(OUT, "nets_odd.pt"))

    # ---- config-1 shape (B=1, 40x40 -> 160x160), one denoiser call at t=50 ----------------
    img_lr, img_lr_up, g = gen_inputs(1, 40, 40, hp["sr_scale"], 13)
    x = torch.randn((1, 3, 160, 160), generator=g)
    t = torch.tensor([50], dtype=torch.long)
    with torch.no_grad():
        rrdb_out, feas = diff.rrdb(img_lr, True)
        eps = diff.denoise_fn(x, t, feas, img_lr_up)
    torch.save({"seed_inputs": 13, "t": t, "eps": eps.half(), "rrdb_out": rrdb_out.half(),
                "eps_norm": float(eps.norm()), "rrdb_out_norm": float(rrdb_out.norm())},
               os.path.join(OUT, "nets_c1.pt"))

    # ---- p_sample arithmetic (row P) -------------------------------------------------------
    g = torch.Generator().manual_seed(17)
    x = torch.randn((4, 3, 8, 12), generator=g)
    e = torch.randn((4, 3, 8, 12), generator=g)
    nz = torch.randn((4, 3, 8, 12), generator=g)
    t = torch.tensor([0, 1, 50, 99], dtype=torch.long)
    with ref_loader.inject_noise(ref, nz[None]):
        with torch.no_grad():
            xp, x0 = diff.p_sample(x, t, None, None, noise_pred=e)
    torch.save({"x": x, "eps": e, "noise": nz, "t": t, "x_prev": xp, "x0": x0,
                "res2img": diff.res2img(xp, x), "img2res": diff.img2res(xp, x)},
               os.path.join(OUT, "p_sample.pt"))

    # ---- full sampler, T=4, B=2, 8x8 -> 32x32, injected noise (rows S, F) -------------------
    ref, diff4 = ref_loader.build_reference(hp, timesteps=4, seed=SEED)
    img_lr, img_lr_up, g = gen_inputs(2, 8, 8, hp["sr_scale"], 19)
    noises = torch.randn((5, 2, 3, 32, 32), generator=g)
    rec = {"eps": []}
    orig = diff4.denoise_fn.forward

    def hooked(*a, **k):
        o = orig(*a, **k)
        rec["eps"].append(o.clone())
        return o
    diff4.denoise_fn.forward = hooked
    with ref_loader.inject_noise(ref, noises):
        with torch.no_grad():
            img, rrdb_out = diff4.sample(img_lr, img_lr_up, tuple(img_lr_up.shape))
    diff4.denoise_fn.forward = orig
    torch.save({"img_lr": img_lr, "img_lr_up": img_lr_up, "noises": noises, "img": img, "rrdb_out": rrdb_out,
                "eps": torch.stack(rec["eps"])}, os.path.join(OUT, "sample_T4.pt"))
    for fn in sorted(os.listdir(OUT)):
        print("%-24s %8d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
