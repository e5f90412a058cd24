"""SURVEY 8f "next" rows: q_sample / interpolate / res=False start and the driver's image handling.
CPU tests pin the oracle restatement against tests/golden/next_rows.pt (generated from the unmodified
reference by oracle/make_golden_next.py); GPU tests compare the CUDA path, through the public API -> C ABI,
with those fixtures and with the oracle."""
import numpy as np
import pytest
import torch

from conftest import build_ours, golden, gen_inputs
from oracle import srdiff_oracle as O

DEV = "cuda:0"


# ------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_q_sample_bicubic_merge_uint8_match_fixture(hp):
    g = golden("next_rows.pt")
    s = O.make_schedule(100, hp)
    q = g["q_sample"]
    out = O.q_sample(s, q["x"], q["t"], q["noise"])
    assert torch.equal(out, q["out"]) and torch.equal(out[3], q["x"][3])          # t < 0: unchanged
    b = g["bicubic"]
    assert torch.equal(O.bicubic_up(b["x"], b["scale"]), b["out"])
    m = g["merger"]
    assert np.array_equal(O.merge_tiles(m["tiles"].numpy(), m["nx"], m["ny"]), m["merged"].numpy())
    u = g["to_uint8"]
    assert torch.equal(O.to_uint8_hwc(u["x"])[0], u["u8"])


def test_oracle_interpolate_and_nores_sample_match_fixture(hp, cpu_state):
    g = golden("next_rows.pt")
    s = O.make_schedule(4, hp)
    it = g["interpolate"]
    with torch.no_grad():
        img = O.interpolate(cpu_state, s, hp, it["x1"], it["x2"], it["img_lr"], it["img_lr_up"], it["noises"],
                            it["t"], it["lam"])
    assert O.psnr(img, it["img"]) > 100.0
    nr = g["sample_nores"]
    with torch.no_grad():
        img, _ = O.sample_nores(cpu_state, s, dict(hp, res=False), nr["img_lr"], nr["img_lr_up"], nr["noises"])
    assert O.psnr(img, nr["img"]) > 100.0


def test_oracle_normalize_roundtrip():
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (2, 5, 7, 3), generator=g, dtype=torch.uint8)
    x = O.normalize_u8(u8)
    assert x.shape == (2, 3, 5, 7) and float(x.min()) >= -1 and float(x.max()) <= 1
    assert int((O.to_uint8_hwc(x).int() - u8.int()).abs().max()) <= 1              # display conversion truncates
    y = O.normalize_u8(u8, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    assert float(y[0, 0, 0, 0]) == pytest.approx((float(u8[0, 0, 0, 0]) / 255 - 0.485) / 0.229, rel=1e-6)


# ------------------------------------------------------------------------------------------ GPU
class _inject:
    """Feed our sampler the fixture's noise: torch.randn_like (q_sample) and model.noise_like (p_sample)."""

    def __init__(self, M, noises):
        self.M, self.noises, self.i = M, noises, 0

    def _next(self, dev):
        n = self.noises[self.i]
        self.i += 1
        return n.clone().to(dev)

    def __enter__(self):
        self.old_like, self.old_nl = torch.randn_like, self.M.noise_like
        torch.randn_like = lambda x, **k: self._next(x.device)
        self.M.noise_like = lambda shape, device, repeat=False: self._next(device)
        return self

    def __exit__(self, *exc):
        torch.randn_like, self.M.noise_like = self.old_like, self.old_nl
        return False


@pytest.mark.gpu
def test_q_sample_bit_exact(hp):
    M, diff = build_ours(hp, device=DEV)
    q = golden("next_rows.pt")["q_sample"]
    out = diff.q_sample(q["x"].to(DEV), q["t"].to(DEV), q["noise"].to(DEV))
    assert torch.equal(out.cpu(), q["out"])


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_interpolate_matches_reference_fixture(hp, prec):
    M, diff = build_ours(hp, timesteps=4, device=DEV)
    M.runtime.precision, M.runtime.noise = prec, "torch"
    it = golden("next_rows.pt")["interpolate"]
    try:
        with _inject(M, it["noises"]):
            img = diff.interpolate(it["x1"].to(DEV), it["x2"].to(DEV), it["img_lr"].to(DEV), it["img_lr_up"].to(DEV),
                                   t=it["t"], lam=it["lam"])
        assert O.psnr(img.cpu(), it["img"]) >= (80.0 if prec == "fp32" else 40.0)
    finally:
        M.runtime.precision = "bf16"


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_sample_without_res_matches_reference_fixture(hp, prec):
    M, diff = build_ours(dict(hp, res=False), timesteps=4, device=DEV)
    M.runtime.precision, M.runtime.noise = prec, "torch"
    nr = golden("next_rows.pt")["sample_nores"]
    try:
        with _inject(M, nr["noises"]):
            img, _ = diff.sample(nr["img_lr"].to(DEV), nr["img_lr_up"].to(DEV), tuple(nr["img_lr_up"].shape))
        assert O.psnr(img.cpu(), nr["img"]) >= (80.0 if prec == "fp32" else 40.0)
    finally:
        M.runtime.precision = "bf16"
        M.hparams["res"] = True


@pytest.mark.gpu
def test_bicubic_upsample_matches_torch_reference():
    from dgm_img_super_resolution_b200 import pipeline as P
    b = golden("next_rows.pt")["bicubic"]
    out = P.bicubic_upsample(b["x"].to(DEV), b["scale"]).cpu()
    assert out.shape == b["out"].shape
    assert float((out - b["out"]).abs().max()) < 2e-6           # fp32, FMA contraction vs ATen's order
    g = torch.Generator().manual_seed(1)
    x = torch.rand((3, 3, 40, 40), generator=g) * 2 - 1         # the bench's LR patch size
    assert float((P.bicubic_upsample(x.to(DEV), 4).cpu() - O.bicubic_up(x, 4)).abs().max()) < 2e-6
    x1 = torch.rand((1, 1, 1, 9), generator=g)                  # degenerate height: scale 0 along H
    assert float((P.bicubic_upsample(x1.to(DEV), 2).cpu() - O.bicubic_up(x1, 2)).abs().max()) < 2e-6


@pytest.mark.gpu
def test_uint8_conversions_bit_exact():
    from dgm_img_super_resolution_b200 import pipeline as P
    u = golden("next_rows.pt")["to_uint8"]
    assert torch.equal(P.to_uint8(u["x"].to(DEV)).cpu()[0], u["u8"])
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, 3, 33, 17), generator=g)
    assert torch.equal(P.to_uint8(x.to(DEV)).cpu(), O.to_uint8_hwc(x))
    u8 = torch.randint(0, 256, (2, 9, 11, 3), generator=g, dtype=torch.uint8)
    assert torch.equal(P.normalize_u8(u8.to(DEV)).cpu(), O.normalize_u8(u8))
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    assert float((P.normalize_u8(u8.to(DEV), mean, std).cpu() - O.normalize_u8(u8, mean, std)).abs().max()) < 1e-6
    with pytest.raises(Exception):
        P.to_uint8(x)                                           # CPU tensor: no fallback


@pytest.mark.gpu
def test_tiles_split_merge_match_imgmerger():
    from dgm_img_super_resolution_b200 import pipeline as P
    m = golden("next_rows.pt")["merger"]
    merged = P.merge_tiles(m["tiles"].to(DEV), m["nx"], m["ny"]).cpu()
    assert torch.equal(merged[0], m["merged"])
    tiles, grid = P.split_tiles(merged.to(DEV), 4, 5)
    assert grid == (2, 3) and torch.equal(tiles.cpu(), m["tiles"])
    calls = []
    mg = P.ImgMerger(lambda a: calls.append(a) or a, device=DEV)
    res = None
    for xx in range(2):
        for yy in range(3):
            res = mg.push([m["tiles"][xx * 3 + yy].numpy()], (xx, yy), (2, 3))
    assert np.array_equal(res, m["merged"].numpy()) and len(calls) == 1
    with pytest.raises(ValueError):
        P.split_tiles(merged.to(DEV), 3, 5)


@pytest.mark.gpu
def test_enhance_pipeline_tiled_equals_batched(hp):
    """configs[2]-style tiling: the tiled run equals sampling the tiles as an explicit batch (same Philox keys)."""
    from dgm_img_super_resolution_b200 import pipeline as P
    M, diff = build_ours(hp, timesteps=3, device=DEV)
    M.runtime.noise, M.runtime.precision = "philox", "bf16"
    try:
        img_lr, _, _ = gen_inputs(1, 16, 24, hp["sr_scale"], 9)
        M.runtime.seed_counter = 7
        out = P.enhance(diff, img_lr.to(DEV), tile=(8, 8))
        assert out.shape == (1, 64, 96, 3) and out.dtype == torch.uint8
        tiles, grid = P.split_tiles(img_lr.to(DEV), 8, 8)
        M.runtime.seed_counter = 7
        ref = diff.sample(tiles, P.bicubic_upsample(tiles, 4), (6, 3, 32, 32))[0]
        assert torch.equal(out, P.to_uint8(P.merge_tiles(ref, *grid)))
    finally:
        M.runtime.noise = "torch"
