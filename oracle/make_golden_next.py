"""TEST INFRASTRUCTURE ONLY -- golden vectors for the SURVEY 8f "next" rows (q_sample, interpolate, the
res=False start, bicubic upsample, ImgMerger), produced by running the UNMODIFIED reference.

    python -m oracle.make_golden_next      ->  tests/golden/next_rows.pt
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import OUT, SEED, gen_inputs  # noqa: E402


class inject_all(ref_loader.inject_noise):
    """inject_noise that also feeds torch.randn_like (q_sample's default noise :1042)."""

    def __enter__(self):
        super().__enter__()
        self._old_like = torch.randn_like
        torch.randn_like = lambda x, **k: self._next().to(x.device)
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._old_like
        return super().__exit__(*exc)


def main():
    hp = ref_loader.load_hparams()
    out = {}
    # ---- q_sample :1041-1048 (T = 100 schedule), including t < 0 ---------------------------
    ref, diff = ref_loader.build_reference(hp, seed=SEED)
    g = torch.Generator().manual_seed(23)
    x = torch.randn((4, 3, 6, 10), generator=g)
    nz = torch.randn((4, 3, 6, 10), generator=g)
    t = torch.tensor([0, 37, 99, -1], dtype=torch.long)
    out["q_sample"] = {"x": x, "noise": nz, "t": t, "out": diff.q_sample(x, t, nz)}

    # ---- interpolate :1093-1116, T = 4, t = 3, lam = 0.3 -----------------------------------
    ref, diff4 = ref_loader.build_reference(hp, timesteps=4, seed=SEED)
    img_lr, img_lr_up, g = gen_inputs(2, 8, 8, hp["sr_scale"], 29)
    x1 = torch.rand((2, 3, 32, 32), generator=g) * 2 - 1
    x2 = torch.rand((2, 3, 32, 32), generator=g) * 2 - 1
    noises = torch.randn((5, 2, 3, 32, 32), generator=g)        # 2 x q_sample, then 3 x p_sample
    ref.tqdm = lambda it, **k: it
    with inject_all(ref, noises):
        with torch.no_grad():
            img = diff4.interpolate(x1, x2, img_lr, img_lr_up, t=3, lam=0.3)
    out["interpolate"] = {"img_lr": img_lr, "img_lr_up": img_lr_up, "x1": x1, "x2": x2, "noises": noises,
                          "t": 3, "lam": 0.3, "img": img}

    # ---- sample() with hparams['res'] = False :1066-1068 ------------------------------------
    hp2 = dict(hp, res=False)
    ref, diff4n = ref_loader.build_reference(hp2, timesteps=4, seed=SEED)
    noises = torch.randn((5, 2, 3, 32, 32), generator=g)        # q_sample start, then 4 x p_sample
    with inject_all(ref, noises):
        with torch.no_grad():
            img, rrdb_out = diff4n.sample(img_lr, img_lr_up, tuple(img_lr_up.shape))
    out["sample_nores"] = {"img_lr": img_lr, "img_lr_up": img_lr_up, "noises": noises, "img": img}
    ref.hparams.update(hp)

    # ---- bicubic x4, align_corners=True (st_app1.py:1593-1597) ------------------------------
    g = torch.Generator().manual_seed(31)
    xin = torch.rand((2, 3, 5, 7), generator=g) * 2 - 1
    out["bicubic"] = {"x": xin, "scale": 4,
                      "out": torch.nn.functional.interpolate(xin, scale_factor=4, mode="bicubic", align_corners=True)}

    # ---- ImgMerger :121-145: 2 x 3 grid of [C, 4, 5] tiles ----------------------------------
    tiles = (torch.rand((6, 3, 4, 5), generator=g) * 255).numpy()
    merger = ref.ImgMerger(lambda a: a)
    merged = None
    for xx in range(2):
        for yy in range(3):
            r = merger.push([tiles[xx * 3 + yy]], (xx, yy), (2, 3))
            merged = r if r is not None else merged
    out["merger"] = {"tiles": torch.from_numpy(tiles), "nx": 2, "ny": 3, "merged": torch.from_numpy(np.asarray(merged))}

    # ---- display conversion st_app1.py:1682-1684, 1710-1711 ---------------------------------
    e = torch.randn((1, 3, 6, 5), generator=g) * 0.8
    v = e.squeeze(0).clamp(-1, 1)
    v = ((v + 1) / 2).permute(1, 2, 0).numpy()
    out["to_uint8"] = {"x": e, "u8": torch.from_numpy(np.clip(v * 255, 0, 255).astype(np.uint8))}
    torch.save(out, os.path.join(OUT, "next_rows.pt"))
    print("next_rows.pt %d bytes" % os.path.getsize(os.path.join(OUT, "next_rows.pt")))


if __name__ == "__main__":
    main()
