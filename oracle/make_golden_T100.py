"""TEST INFRASTRUCTURE ONLY -- the reference's FULL configs[0] trajectory: one 40x40 -> 160x160 image through all
T = 100 reverse-diffusion steps (`srdiff/config.yaml`), with the noise of a seeded CPU generator injected, produced by
running the UNMODIFIED reference.  Only the final image is stored (fp32, 300 KB); the tests regenerate the inputs and
the 101 noise tensors from the same seeds.

    python -m oracle.make_golden_T100      ->  tests/golden/sample_T100.pt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import OUT, SEED, gen_inputs  # noqa: E402

SEED_INPUTS, SEED_NOISE = 83, 89


def make_noises(T, shape):
    g = torch.Generator().manual_seed(SEED_NOISE)
    return torch.randn((T + 1,) + tuple(shape), generator=g)


def main():
    hp = ref_loader.load_hparams()
    ref, diff = ref_loader.build_reference(hp, seed=SEED)           # timesteps = hp['timesteps'] = 100
    T = diff.num_timesteps
    img_lr, img_lr_up, _ = gen_inputs(1, 40, 40, hp["sr_scale"], SEED_INPUTS)
    noises = make_noises(T, img_lr_up.shape)
    with ref_loader.inject_noise(ref, noises):
        with torch.no_grad():
            img, rrdb_out = diff.sample(img_lr, img_lr_up, tuple(img_lr_up.shape))
    torch.save({"T": T, "seed_inputs": SEED_INPUTS, "seed_noise": SEED_NOISE, "img": img,
                "img_abs_max": float(img.abs().max()), "img_norm": float(img.norm())},
               os.path.join(OUT, "sample_T100.pt"))
    print("sample_T100.pt %d bytes, |img|max %.3f" % (os.path.getsize(os.path.join(OUT, "sample_T100.pt")),
                                                      float(img.abs().max())))


if __name__ == "__main__":
    main()
