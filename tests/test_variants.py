"""sr_scale = 8, up_input, weight norm, the GroupNorm `Block`, stand-alone
`ResnetBlock(x, time_emb, cond)`, checkpoint files, torch SSIM and the uint8 PSNR.
CPU tests pin the oracle restatement against tests/golden/variants.pt (generated from the unmodified reference by
oracle/make_golden_variants.py) and cover the host logic; GPU tests compare the CUDA path, through the public API
-> C ABI, with those fixtures."""
import os

import numpy as np
import pytest
import torch

from conftest import SEED, build_ours, golden
from oracle import ref_loader
from oracle import srdiff_oracle as O
from oracle.make_golden_variants import perturb_

DEV = "cuda:0"
TOL = {"fp32": 1e-4, "bf16": 1e-2}          # rel-L2 of activations (north_star tolerances)


def _sd(diff):
    return {k: v.detach().clone().cpu() for k, v in diff.state_dict().items()}


class _inject:
    def __init__(self, M, noises):
        self.M, self.noises, self.i = M, noises, 0

    def _next(self, dev):
        n = self.noises[self.i]
        self.i += 1
        return n.clone().to(dev)

    def __enter__(self):
        self.old_randn, self.old_nl = torch.randn, self.M.noise_like
        torch.randn = lambda *a, **k: self._next(k.get("device", "cpu"))
        self.M.noise_like = lambda shape, device, repeat=False: self._next(device)
        return self

    def __exit__(self, *exc):
        torch.randn, self.M.noise_like = self.old_randn, self.old_nl
        return False


def _blocks(M):
    """The four stand-alone blocks of the fixture, seeded and perturbed like oracle/make_golden_variants.py."""
    torch.manual_seed(SEED)
    blk = M.Block(8, 16, groups=8)
    perturb_(blk)
    torch.manual_seed(SEED + 1)
    rb_gn = M.ResnetBlock(8, 16, time_emb_dim=12, groups=4)
    perturb_(rb_gn)
    torch.manual_seed(SEED + 2)
    rb = M.ResnetBlock(8, 16, time_emb_dim=12, groups=0)
    torch.manual_seed(SEED + 3)
    rb_same = M.ResnetBlock(16, 16, time_emb_dim=12, groups=2)
    perturb_(rb_same)
    return blk, rb_gn, rb, rb_same


# ------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_sr8_and_up_input_match_fixture(hp):
    g = golden("variants.pt")
    for key, hpv in (("sr8", dict(hp, sr_scale=8)), ("up_input", dict(hp, up_input=True))):
        M, diff = build_ours(hpv, timesteps=2)
        f = g[key]
        with torch.no_grad():
            img, rrdb_out = O.sample(_sd(diff), O.make_schedule(2, hpv), hpv, f["img_lr"], f["img_lr_up"], f["noises"])
        assert O.psnr(img, f["img"]) > 100.0, key
        if "rrdb_out" in f:
            assert O.rel_l2(rrdb_out, f["rrdb_out"]) < 5e-6
    M.hparams.update(hp)


def test_oracle_weight_norm_matches_fixture_and_state_dict_keys(hp):
    f = golden("variants.pt")["use_wn"]
    hpw = dict(hp, use_wn=True)
    M, diff = build_ours(hpw, timesteps=2)
    try:
        perturb_(diff.denoise_fn)
        sd = _sd(diff)
        n_g = sum(1 for k in sd if k.endswith("weight_g"))
        assert n_g == f["n_wn"] and "denoise_fn.cond_proj.weight" in sd            # transposed conv: no weight norm
        assert "denoise_fn.downs.0.0.block1.block.1.weight" not in sd
        with torch.no_grad():
            _, feas = O.rrdb_forward(sd, f["img_lr"], hpw)
            eps = O.unet_forward(sd, f["x"], f["t"], feas, hpw)
        assert O.rel_l2(eps, f["eps"]) < 5e-6
        assert O.rel_l2(f["eps_fast"], f["eps"]) < 5e-6
        # folding in the engine == removing the reparametrisation (make_generation_fast_ :849-856)
        from dgm_img_super_resolution_b200 import engine as E
        conv = diff.denoise_fn.downs[1][0].block1.block[1]
        w_eff = E.eff_weight(conv).clone()
        assert E.eff_weight(conv) is E.eff_weight(conv)                             # cached until g / v change
        diff.denoise_fn.make_generation_fast_()
        assert not any(k.endswith("weight_g") for k in diff.state_dict())
        assert torch.allclose(conv.weight, w_eff, rtol=1e-6, atol=1e-7)
        assert E.eff_weight(conv) is conv.weight
    finally:
        M.hparams.update(hp)


def test_oracle_linear_attention_matches_fixture(hp):
    f = golden("variants.pt")["use_attn"]
    hpa = dict(hp, use_attn=True)
    M, diff = build_ours(hpa, timesteps=2)
    try:
        with torch.no_grad():
            diff.denoise_fn.mid_attn.fn.g.fill_(400.0)
            sd = _sd(diff)
            assert sd["denoise_fn.mid_attn.fn.fn.to_qkv.weight"].shape == (384, 256, 1, 1)
            assert "denoise_fn.mid_attn.fn.fn.to_qkv.bias" not in sd
            assert O.rel_l2(O.linear_attention(sd, "denoise_fn.mid_attn.fn.fn.", f["xa"]), f["attn"]) < 5e-6
            _, feas = O.rrdb_forward(sd, f["img_lr"], hpa)
            assert O.rel_l2(O.unet_forward(sd, f["x"], f["t"], feas, hpa), f["eps"]) < 5e-6
            sd["denoise_fn.mid_attn.fn.g"] = torch.zeros(1)            # the gain matters
            assert O.rel_l2(O.unet_forward(sd, f["x"], f["t"], feas, hpa), f["eps"]) > 2e-2
    finally:
        M.hparams.update(hp)


def test_oracle_groupnorm_blocks_match_fixture(hp):
    import dgm_img_super_resolution_b200.model as M
    f = golden("variants.pt")["gn"]
    blk, rb_gn, rb, rb_same = _blocks(M)
    with torch.no_grad():
        sd = {k: v for k, v in blk.state_dict().items()}
        assert O.rel_l2(O._block(sd, "", f["x"], groups=8), f["block"]) < 5e-6
        sd = dict(rb_gn.state_dict())
        assert O.rel_l2(O.resnet_block(sd, "", f["x"], f["time_emb"], groups=4), f["resnet_gn"]) < 5e-6
        assert O.rel_l2(O.resnet_block(sd, "", f["x"], groups=4), f["resnet_gn_plain"]) < 5e-6
        sd = dict(rb.state_dict())
        assert O.rel_l2(O.resnet_block(sd, "", f["x"], f["time_emb"], f["cond"]), f["resnet"]) < 5e-6
        assert O.rel_l2(O.resnet_block(sd, "", f["x"], f["time_emb"]), f["resnet_te"]) < 5e-6
        sd = dict(rb_same.state_dict())
        assert O.rel_l2(O.resnet_block(sd, "", f["cond"], f["time_emb"], groups=2), f["resnet_gn_same"]) < 5e-6


def test_oracle_ssim_and_psnr(hp):
    f = golden("variants.pt")["ssim"]
    assert float(O.ssim_torch(f["a"], f["b"])) == pytest.approx(float(f["module"]), abs=1e-6)
    assert torch.allclose(O.ssim_torch(f["a"], f["b"], size_average=False), f["per_image"], atol=1e-6)
    assert float(O.ssim_torch(f["a"], f["b"], map01=False)) == pytest.approx(float(f["functional"]), abs=1e-6)
    assert torch.allclose(O.ssim_torch(f["a"], f["b"], 7, False, map01=False), f["win7"], atol=1e-6)
    a, b = O.to_uint8_round(f["a"]), O.to_uint8_round(f["b"])
    assert a.dtype == np.uint8 and O.psnr_u8(a, a) == float("inf")
    mse = np.mean((a.astype(np.int64) - b.astype(np.int64)) ** 2)
    assert O.psnr_u8(a, b) == pytest.approx(10 * np.log10(255.0 ** 2 / mse), rel=1e-12)


def test_checkpoint_files_round_trip_and_match_the_reference_format(hp, tmp_path):
    M, diff = build_ours(hp, timesteps=4)
    opt = torch.optim.SGD(diff.parameters(), lr=0.1)
    assert not hasattr(M, "save_checkpoint") and not hasattr(M, "load_checkpoint")    # trainer halves: out of scope
    for step in (2000, 300):          # files in the upstream trainer's layout (:1441-1449), written by the test
        torch.save({"global_step": step, "optimizer_states": [opt.state_dict()],
                    "state_dict": {"model": diff.state_dict()}}, str(tmp_path / ("model_ckpt_steps_%d.ckpt" % step)),
                   _use_new_zipfile_serialization=False)
    names = [os.path.basename(p) for p in M.get_all_ckpts(str(tmp_path))]
    assert names == ["model_ckpt_steps_2000.ckpt", "model_ckpt_steps_300.ckpt"]       # newest first
    ckpt, path = M.get_last_checkpoint(str(tmp_path))
    assert path.endswith("model_ckpt_steps_2000.ckpt") and ckpt["global_step"] == 2000
    assert set(ckpt) == {"global_step", "optimizer_states", "state_dict"} and set(ckpt["state_dict"]) == {"model"}
    M2, other = build_ours(hp, timesteps=4, seed=SEED + 9)
    w = "denoise_fn.mid_block1.block1.block.1.weight"
    assert not torch.equal(other.state_dict()[w], diff.state_dict()[w])
    M.load_ckpt(other, str(tmp_path))                                                  # directory: newest file
    assert all(torch.equal(v, other.state_dict()[k]) for k, v in diff.state_dict().items())
    # flat 'model.'-prefixed layout (:1473-1475) and strict=False shape filtering (:1478-1487)
    flat = {"state_dict": {"model." + k: v for k, v in diff.state_dict().items()}}
    flat["state_dict"]["model.denoise_fn.mlp.0.bias"] = torch.zeros(3)
    fp = str(tmp_path / "flat.ckpt")
    torch.save(flat, fp)
    _, third = build_ours(hp, timesteps=4, seed=SEED + 11)
    keep = third.state_dict()["denoise_fn.mlp.0.bias"].clone()
    M.load_ckpt(third, fp, strict=False)
    assert torch.equal(third.state_dict()["denoise_fn.mlp.0.bias"], keep)
    assert torch.equal(third.state_dict()[w], diff.state_dict()[w])
    with pytest.raises(AssertionError):
        M.load_ckpt(third, str(tmp_path / "nothing_here"))
    M.load_ckpt(third, str(tmp_path / "nothing_here"), force=False)                   # prints, does not raise
    if ref_loader.available():        # the reference reads our files and we read its state dict
        ref, rdiff = ref_loader.build_reference(hp, timesteps=4, seed=SEED + 13)
        ref.glob = __import__("glob")                                                  # un-imported upstream
        ref.load_ckpt(rdiff, str(tmp_path))
        assert all(torch.equal(v, rdiff.state_dict()[k]) for k, v in diff.state_dict().items())


def test_initialize_weights_semantics():
    import dgm_img_super_resolution_b200.model as M
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.Linear(5, 7), torch.nn.BatchNorm2d(8))
    torch.manual_seed(5)
    M.initialize_weights(net, scale=0.1)
    torch.manual_seed(5)
    w = torch.empty(8, 4, 3, 3)
    torch.nn.init.kaiming_normal_(w, a=0, mode="fan_in")
    assert torch.equal(net[0].weight.data, w * 0.1) and float(net[0].bias.abs().sum()) == 0
    assert float(net[1].bias.abs().sum()) == 0 and float(net[2].weight.min()) == 1.0


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("key", ["sr8", "up_input"])
def test_sample_variants_match_reference_fixture(hp, prec, key):
    hpv = dict(hp, sr_scale=8) if key == "sr8" else dict(hp, up_input=True)
    M, diff = build_ours(hpv, timesteps=2, device=DEV)
    M.runtime.precision, M.runtime.noise = prec, "torch"
    f = golden("variants.pt")[key]
    try:
        with _inject(M, f["noises"]):
            img, rrdb_out = diff.sample(f["img_lr"].to(DEV), f["img_lr_up"].to(DEV), tuple(f["img_lr_up"].shape))
        assert O.psnr(img.cpu(), f["img"]) >= (80.0 if prec == "fp32" else 40.0)
        if "rrdb_out" in f:
            assert O.rel_l2(rrdb_out.cpu(), f["rrdb_out"]) < TOL[prec]
    finally:
        M.runtime.precision = "bf16"
        M.hparams.update(hp)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_weight_norm_unet_matches_reference_fixture(hp, prec):
    hpw = dict(hp, use_wn=True)
    M, diff = build_ours(hpw, timesteps=2)
    perturb_(diff.denoise_fn)
    diff = diff.to(DEV)
    M.runtime.precision = prec
    f = golden("variants.pt")["use_wn"]
    try:
        _, feas = diff.rrdb(f["img_lr"].to(DEV), True)
        args = (f["x"].to(DEV), f["t"].to(DEV), feas, f["img_lr_up"].to(DEV))
        eps = diff.denoise_fn(*args)
        assert O.rel_l2(eps.cpu(), f["eps"]) < TOL[prec]
        with torch.no_grad():                    # a changed gain must be picked up (no stale packed weights)
            diff.denoise_fn.final_conv[0].block[1].weight_g.mul_(2.0)
        assert O.rel_l2(diff.denoise_fn(*args).cpu(), f["eps"]) > 0.1
        with torch.no_grad():
            diff.denoise_fn.final_conv[0].block[1].weight_g.mul_(0.5)
        diff.denoise_fn.make_generation_fast_()
        assert O.rel_l2(diff.denoise_fn(*args).cpu(), f["eps_fast"]) < TOL[prec]
    finally:
        M.runtime.precision = "bf16"
        M.hparams.update(hp)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_linear_attention_unet_matches_reference_fixture(hp, prec):
    hpa = dict(hp, use_attn=True)
    M, diff = build_ours(hpa, timesteps=2, device=DEV)
    with torch.no_grad():
        diff.denoise_fn.mid_attn.fn.g.fill_(400.0)
    M.runtime.precision = prec
    f = golden("variants.pt")["use_attn"]
    try:
        attn = diff.denoise_fn.mid_attn.fn.fn(f["xa"].to(DEV))
        assert O.rel_l2(attn.cpu(), f["attn"]) < TOL[prec]
        _, feas = diff.rrdb(f["img_lr"].to(DEV), True)
        eps = diff.denoise_fn(f["x"].to(DEV), f["t"].to(DEV), feas, f["img_lr_up"].to(DEV))
        assert O.rel_l2(eps.cpu(), f["eps"]) < TOL[prec]
        with torch.no_grad():                                       # Rezero gain is read when the program is built
            diff.denoise_fn.mid_attn.fn.g.fill_(0.0)
        eps0 = diff.denoise_fn(f["x"].to(DEV), f["t"].to(DEV), feas, f["img_lr_up"].to(DEV))
        assert O.rel_l2(eps0.cpu(), f["eps"]) > 2e-2
    finally:
        M.runtime.precision = "bf16"
        M.hparams.update(hp)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_groupnorm_and_standalone_resnet_blocks_match_reference_fixture(hp, prec):
    import dgm_img_super_resolution_b200.model as M
    f = golden("variants.pt")["gn"]
    blk, rb_gn, rb, rb_same = [m.to(DEV) for m in _blocks(M)]
    x, te, cond = f["x"].to(DEV), f["time_emb"].to(DEV), f["cond"].to(DEV)
    M.runtime.precision = prec
    tol = TOL[prec] * (1.0 if prec == "fp32" else 1.5)      # two bf16 roundings around each normalisation
    try:
        assert O.rel_l2(blk(x).cpu(), f["block"]) < tol
        assert O.rel_l2(rb_gn(x, te).cpu(), f["resnet_gn"]) < tol
        assert O.rel_l2(rb_gn(x).cpu(), f["resnet_gn_plain"]) < tol
        assert O.rel_l2(rb(x, te, cond).cpu(), f["resnet"]) < tol
        assert O.rel_l2(rb(x, te).cpu(), f["resnet_te"]) < tol
        assert O.rel_l2(rb_same(cond, te).cpu(), f["resnet_gn_same"]) < tol
    finally:
        M.runtime.precision = "bf16"


@pytest.mark.gpu
def test_groupnorm_rejects_bad_arguments():
    from dgm_img_super_resolution_b200 import _lib as L
    a = L.SrdGroupNormArgs()
    buf = torch.zeros(64, device=DEV)
    a.src = a.out = buf.data_ptr()
    a.gamma = a.beta = a.stats = buf.data_ptr()
    a.B, a.H, a.W, a.C, a.groups, a.src_ct, a.out_ct = 1, 2, 2, 6, 4, 8, 8
    import ctypes as C
    assert L.lib().srd_groupnorm_act(C.byref(a), None) != 0
    assert b"in-place" in L.lib().srd_last_error()
    out = torch.zeros(256, device=DEV)
    a.out = out.data_ptr()
    assert L.lib().srd_groupnorm_act(C.byref(a), None) != 0
    assert b"multiple of groups" in L.lib().srd_last_error()


@pytest.mark.gpu
def test_ssim_and_psnr_match_reference_fixture():
    from dgm_img_super_resolution_b200 import metrics as Q
    f = golden("variants.pt")["ssim"]
    a, b = f["a"].to(DEV), f["b"].to(DEV)
    assert float(Q.SSIM()(a, b)) == pytest.approx(float(f["module"]), abs=2e-6)
    assert torch.allclose(Q.SSIM(size_average=False)(a, b).cpu(), f["per_image"], atol=2e-6)
    assert float(Q.ssim(a, b)) == pytest.approx(float(f["functional"]), abs=2e-6)
    assert torch.allclose(Q.ssim(a, b, window_size=7, size_average=False).cpu(), f["win7"], atol=2e-6)
    assert float(Q.SSIM()(a, a)) == pytest.approx(1.0, abs=1e-6)
    # uint8 conversion and PSNR: integer work, bit-exact against the numpy restatement
    ua, ub = Q.to_uint8_round(a), Q.to_uint8_round(b)
    assert np.array_equal(ua.cpu().numpy(), O.to_uint8_round(f["a"]))
    half = torch.tensor([-1.0, -0.996078431, 0.0, 0.00392157, 1.0, 1.5, -2.0], device=DEV)      # ties / clipping
    assert np.array_equal(Q.to_uint8_round(half).cpu().numpy(), O.to_uint8_round(half.cpu()))
    m = Q.Measure()
    assert m.psnr(a, b) == pytest.approx(O.psnr_u8(O.to_uint8_round(f["a"]), O.to_uint8_round(f["b"])), rel=1e-12)
    assert Q.psnr_u8(ua, ua) == float("inf")
    g = torch.Generator().manual_seed(7)                                                          # a large case
    big_a = torch.randint(0, 256, (4, 3, 512, 512), generator=g, dtype=torch.uint8)
    big_b = torch.randint(0, 256, (4, 3, 512, 512), generator=g, dtype=torch.uint8)
    assert Q.psnr_u8(big_a.to(DEV), big_b.to(DEV)) == pytest.approx(O.psnr_u8(big_a.numpy(), big_b.numpy()), rel=1e-12)
    with pytest.raises(NotImplementedError):
        m.lpips(a, b)


# ------------------------------------------------------------------------------------------ imresize / Measure
_RESIZE_CASES = [("u8", dict(scale=1 / 4), "u8_quarter"), ("u8", dict(sizes=(9, 31)), "u8_sizes_9x31"),
                 ("u8", dict(scale=2), "u8_x2"), ("u8", dict(scale=1 / 3, method="bilinear"), "u8_bilinear_third"),
                 ("f64", dict(scale=0.5), "f64_half"), ("f64", dict(scale=3), "f64_x3")]


def test_oracle_imresize_matches_fixture_and_host_taps_agree():
    from dgm_img_super_resolution_b200 import metrics as Q
    f = golden("variants.pt")["imresize"]
    for src, kw, key in _RESIZE_CASES:
        got = O.imresize_np(f[src].numpy(), **kw)
        if src == "u8":
            assert np.array_equal(got, f[key].numpy()), key                       # integer work: bit-exact
        else:
            assert np.abs(got - f[key].numpy()).max() < 1e-10, key
    # the product's host-side tap tables (metrics.contributions) against the oracle's per-sample restatement
    for n_in, n_out, scale, kern, okern in ((24, 6, 0.25, Q._cubic, O._k_cubic), (20, 31, 31 / 20, Q._cubic, O._k_cubic),
                                            (24, 8, 1 / 3, Q._triangle, O._k_triangle)):
        w, idx = Q.contributions(n_in, n_out, scale, kern, 4.0)
        taps = O._resize_taps(n_in, n_out, scale, okern)
        assert w.shape[0] == n_out and all(np.allclose(w[o], taps[o][0], rtol=0, atol=1e-15) and
                                           np.array_equal(idx[o], taps[o][1]) for o in range(n_out))


def test_oracle_ssim_uniform_properties():
    g = torch.Generator().manual_seed(71)
    a = torch.randint(0, 256, (20, 18, 3), generator=g, dtype=torch.uint8).numpy()
    b = np.clip(a.astype(np.int64) + torch.randint(-20, 21, (20, 18, 3), generator=g).numpy(), 0, 255).astype(np.uint8)
    assert O.ssim_uniform_u8(a, a) == pytest.approx(1.0, abs=1e-12)
    s = O.ssim_uniform_u8(a, b)
    assert 0.5 < s < 1.0 and s == pytest.approx(O.ssim_uniform_u8(b, a), rel=1e-12)           # symmetric
    assert O.ssim_uniform_u8(a[:, ::-1].copy(), b[:, ::-1].copy()) == pytest.approx(s, rel=1e-12)   # flip-invariant


@pytest.mark.gpu
def test_imresize_matches_reference_fixture():
    from dgm_img_super_resolution_b200 import metrics as Q
    f = golden("variants.pt")["imresize"]
    for src, kw, key in _RESIZE_CASES:
        got = Q.imresize(f[src].to(DEV), **kw).cpu()
        assert got.shape == f[key].shape and got.dtype == f[key].dtype, key
        if src == "u8":
            assert torch.equal(got, f[key]), key
        else:
            assert float((got - f[key]).abs().max()) < 1e-10, key
    g = torch.Generator().manual_seed(73)                                            # st_app-sized image, 2-D input
    big = torch.randint(0, 256, (640, 480), generator=g, dtype=torch.uint8)
    assert np.array_equal(Q.imresize(big.to(DEV), 1 / 4).cpu().numpy(), O.imresize_np(big.numpy(), 1 / 4))
    with pytest.raises(RuntimeError):
        Q.imresize(torch.zeros((4, 4), device=DEV), 2)                               # float32: not a reference dtype


@pytest.mark.gpu
def test_measure_matches_oracle():
    from dgm_img_super_resolution_b200 import metrics as Q
    g = torch.Generator().manual_seed(79)
    hr = torch.rand((3, 64, 48), generator=g) * 2 - 1
    sr = (hr + 0.1 * torch.randn((3, 64, 48), generator=g)).clamp(-1, 1)
    lr = torch.rand((3, 16, 12), generator=g) * 2 - 1
    res = Q.Measure().measure(sr.to(DEV), hr.to(DEV), lr.to(DEV), 4)
    A, B_, LR = [O.to_uint8_round(v).transpose(1, 2, 0) for v in (sr, hr, lr)]
    assert set(res) == {"psnr", "ssim", "lr_psnr"}
    assert res["psnr"] == pytest.approx(O.psnr_u8(A, B_), rel=1e-12)
    assert res["lr_psnr"] == pytest.approx(O.psnr_u8(O.imresize_np(A, 1 / 4), LR), rel=1e-12)
    assert res["ssim"] == pytest.approx(O.ssim_uniform_u8(A, B_), rel=1e-9)
    a8 = torch.from_numpy(np.ascontiguousarray(A)).to(DEV)
    assert Q.ssim_uniform_u8(a8, a8) == pytest.approx(1.0, abs=1e-12)
    res2 = Q.Measure(lpips_fn=lambda x, y: 0.25).measure(sr.to(DEV), hr.to(DEV), lr.to(DEV), 4)
    assert res2["lpips"] == 0.25
    with pytest.raises(NotImplementedError):
        Q.Measure().lpips(a8, a8)
