"""TEST INFRASTRUCTURE ONLY -- golden vectors for the SURVEY 8f-3 / 8f-4 rows (sr_scale = 8, up_input, weight
norm, the GroupNorm `Block`, stand-alone `ResnetBlock(x, time_emb, cond)`, torch SSIM), produced by running the
UNMODIFIED reference.

    python -m oracle.make_golden_variants      ->  tests/golden/variants.pt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import OUT, SEED, gen_inputs  # noqa: E402


def perturb_(module):
    """Deterministic, RNG-free perturbation of weight-norm gains and GroupNorm affines (at their default
    initialisation both are the identity, which would not test anything).  The tests apply the same function
    to this package's modules."""
    for name, p in module.named_parameters():
        if name.endswith("weight_g") or "block.2." in name:
            i = torch.arange(p.numel(), dtype=torch.float32).reshape(p.shape)
            with torch.no_grad():
                if name.endswith("bias"):
                    p.copy_(0.1 * torch.cos(0.7 * i))
                else:
                    p.mul_(1.0 + 0.25 * torch.sin(0.37 * i + len(name)))


def run_sample(ref, diff, img_lr, img_lr_up, noises):
    with ref_loader.inject_noise(ref, noises):
        with torch.no_grad():
            img, rrdb_out = diff.sample(img_lr, img_lr_up, tuple(img_lr_up.shape))
    return img, rrdb_out


def main():
    hp = ref_loader.load_hparams()
    out = {}
    # ---- sr_scale = 8 (:720-721, 740-741, cond_proj 16/8/4 :759-761): T = 2 sample on a 6x6 image ------------
    hp8 = dict(hp, sr_scale=8)
    ref, diff = ref_loader.build_reference(hp8, timesteps=2, seed=SEED)
    img_lr, img_lr_up, g = gen_inputs(1, 6, 6, 8, 41)
    noises = torch.randn((3, 1, 3, 48, 48), generator=g)
    img, rrdb_out = run_sample(ref, diff, img_lr, img_lr_up, noises)
    out["sr8"] = {"img_lr": img_lr, "img_lr_up": img_lr_up, "noises": noises, "img": img, "rrdb_out": rrdb_out}

    # ---- up_input (:803-806, 831-832): T = 2 sample -------------------------------------------------------
    hpu = dict(hp, up_input=True)
    ref, diff = ref_loader.build_reference(hpu, timesteps=2, seed=SEED)
    img_lr, img_lr_up, g = gen_inputs(2, 8, 8, hp["sr_scale"], 43)
    noises = torch.randn((3, 2, 3, 32, 32), generator=g)
    img, _ = run_sample(ref, diff, img_lr, img_lr_up, noises)
    out["up_input"] = {"img_lr": img_lr, "img_lr_up": img_lr_up, "noises": noises, "img": img}

    # ---- use_wn (:808-814): perturbed gains, Unet eps + the same after make_generation_fast_ ---------------
    hpw = dict(hp, use_wn=True)
    ref, diff = ref_loader.build_reference(hpw, timesteps=2, seed=SEED)
    perturb_(diff.denoise_fn)
    img_lr, img_lr_up, g = gen_inputs(2, 8, 8, hp["sr_scale"], 47)
    x = torch.randn((2, 3, 32, 32), generator=g)
    t = torch.tensor([1, 0], dtype=torch.long)
    with torch.no_grad():
        _, feas = diff.rrdb(img_lr, True)
        eps = diff.denoise_fn(x, t, feas, img_lr_up)
        diff.denoise_fn.make_generation_fast_()
        eps_fast = diff.denoise_fn(x, t, feas, img_lr_up)
    assert not any(k.endswith("weight_g") for k in diff.state_dict())
    out["use_wn"] = {"img_lr": img_lr, "img_lr_up": img_lr_up, "x": x, "t": t, "eps": eps, "eps_fast": eps_fast,
                     "n_wn": sum(1 for m in diff.denoise_fn.modules() if isinstance(m, torch.nn.Conv2d))}
    ref.hparams.update(hp)

    # ---- use_attn (:506-522, 788-789, 836-837).  The reference never imports einops' `rearrange`: the
    # missing name is supplied here, nothing else is touched.  Rezero's gain starts at 0; it is set to 400 so
    # that the branch moves eps by several per cent.
    import einops
    hpa = dict(hp, use_attn=True)
    ref, diff = ref_loader.build_reference(hpa, timesteps=2, seed=SEED)
    ref.rearrange = einops.rearrange
    with torch.no_grad():
        diff.denoise_fn.mid_attn.fn.g.fill_(400.0)
    img_lr, img_lr_up, g = gen_inputs(2, 8, 8, hp["sr_scale"], 61)
    x = torch.randn((2, 3, 32, 32), generator=g)
    xa = torch.randn((2, 256, 6, 5), generator=g)
    t = torch.tensor([0, 1], dtype=torch.long)
    with torch.no_grad():
        _, feas = diff.rrdb(img_lr, True)
        out["use_attn"] = {"img_lr": img_lr, "img_lr_up": img_lr_up, "x": x, "t": t,
                           "eps": diff.denoise_fn(x, t, feas, img_lr_up),
                           "xa": xa, "attn": diff.denoise_fn.mid_attn.fn.fn(xa)}
    ref.hparams.update(hp)

    # ---- GroupNorm Block / stand-alone ResnetBlock (:439-480) ----------------------------------------------
    g = torch.Generator().manual_seed(53)
    x = torch.randn((2, 8, 10, 12), generator=g)
    te = torch.randn((2, 12), generator=g)
    cond = torch.randn((2, 16, 10, 12), generator=g)
    torch.manual_seed(SEED)
    blk = ref.Block(8, 16, groups=8)
    perturb_(blk)
    torch.manual_seed(SEED + 1)
    rb_gn = ref.ResnetBlock(8, 16, time_emb_dim=12, groups=4)
    perturb_(rb_gn)
    torch.manual_seed(SEED + 2)
    rb = ref.ResnetBlock(8, 16, time_emb_dim=12, groups=0)
    torch.manual_seed(SEED + 3)
    rb_same = ref.ResnetBlock(16, 16, time_emb_dim=12, groups=2)
    perturb_(rb_same)
    with torch.no_grad():
        out["gn"] = {"x": x, "time_emb": te, "cond": cond,
                     "block": blk(x),
                     "resnet_gn": rb_gn(x, te), "resnet_gn_plain": rb_gn(x),
                     "resnet": rb(x, te, cond), "resnet_te": rb(x, te),
                     "resnet_gn_same": rb_same(cond, te)}

    # ---- torch SSIM (:146-222) ------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(59)
    a = torch.rand((2, 3, 20, 24), generator=g) * 2 - 1
    b = (a + 0.2 * torch.randn((2, 3, 20, 24), generator=g)).clamp(-1, 1)
    with torch.no_grad():
        out["ssim"] = {"a": a, "b": b, "module": ref.SSIM()(a, b), "per_image": ref.SSIM(size_average=False)(a, b),
                       "functional": ref.ssim(a, b), "win7": ref.ssim(a, b, window_size=7, size_average=False)}
    # ---- imresize (:1162-1306), uint8 and float64, shrinking (antialiased) and enlarging -----------------------
    g = torch.Generator().manual_seed(67)
    u8 = torch.randint(0, 256, (24, 20, 3), generator=g, dtype=torch.uint8).numpy()
    f64 = torch.rand((12, 10), generator=g, dtype=torch.float64).numpy() * 255
    out["imresize"] = {
        "u8": torch.from_numpy(u8), "f64": torch.from_numpy(f64),
        "u8_quarter": torch.from_numpy(ref.imresize(u8, 1 / 4)),
        "u8_sizes_9x31": torch.from_numpy(ref.imresize(u8, sizes=(9, 31))),
        "u8_x2": torch.from_numpy(ref.imresize(u8, 2)),
        "u8_bilinear_third": torch.from_numpy(ref.imresize(u8, 1 / 3, method="bilinear")),
        "f64_half": torch.from_numpy(ref.imresize(f64, 0.5)),
        "f64_x3": torch.from_numpy(ref.imresize(f64, 3)),
    }
    torch.save(out, os.path.join(OUT, "variants.pt"))
    print("variants.pt %d bytes" % os.path.getsize(os.path.join(OUT, "variants.pt")))


if __name__ == "__main__":
    main()
