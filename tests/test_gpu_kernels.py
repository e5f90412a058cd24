"""GPU parity tests (-m gpu): every CUDA kernel, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Tolerances (north_star): fp32 validation mode rel-L2 <= 1e-4 (we
assert 2e-5 per layer), bf16 rel-L2 <= 1e-2, integer/byte-exact where the arithmetic is
order-identical (layout converters, res2img, the p_sample update up to expf round-off)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import build_ours, gen_inputs, golden
from oracle import srdiff_oracle as O

pytestmark = pytest.mark.gpu
TOL = {"fp32": 2e-5, "bf16_direct": 1e-2, "bf16": 1e-2}
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ours(hp):
    M, diff = build_ours(hp, device=DEV)
    return M, diff


@pytest.fixture()
def precision(ours):
    M, _ = ours
    old = M.runtime.precision
    yield M.runtime
    M.runtime.precision = old


def test_extension_loaded_and_device_is_blackwell():
    from dgm_img_super_resolution_b200 import _lib
    assert _lib.lib().srd_version() >= 100
    assert torch.cuda.get_device_capability(0)[0] == 10


def test_layout_roundtrip_and_rings():
    from dgm_img_super_resolution_b200 import engine as E, _lib as L
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 5, 6, 9), generator=g).to(DEV)
    for prec_name, exact in (("fp32", True), ("bf16", False)):
        for border in (L.BORDER_KEEP, L.BORDER_REFLECT):
            prec = E.Precision(prec_name)
            a = E.Act(2, 6, 9, 8, prec, DEV)
            L.check(L.lib().srd_nchw_to_act(x.data_ptr(), 2, 5, 6, 9, 2.0, 0.5, a.ptr, a.ct, prec.dtype, border,
                                            E._stream()))
            want = x * 2.0 + 0.5
            wantp = F.pad(want, (1, 1, 1, 1), mode="reflect" if border else "constant").permute(0, 2, 3, 1)
            got = a.t.float()
            ref = wantp if exact else wantp.bfloat16().float()
            assert torch.equal(got[..., :5], ref)
            assert float(got[..., 5:].abs().max()) == 0.0
            back = torch.empty_like(x)
            L.check(L.lib().srd_act_to_nchw(a.ptr, a.ct, 0, prec.dtype, 2, 5, 6, 9, back.data_ptr(), E._stream()))
            assert torch.equal(back, ref[:, 1:-1, 1:-1].permute(0, 3, 1, 2))


def test_mish_and_res2img(ours, hp):
    M, diff = ours
    g = torch.Generator().manual_seed(4)
    x = torch.randn((3, 3, 8, 12), generator=g) * 6
    x[0, 0, 0, :4] = torch.tensor([25.0, -30.0, 20.0, 0.0])
    up = torch.randn((3, 3, 8, 12), generator=g)
    assert torch.allclose(M.Mish()(x.to(DEV)).cpu(), O.mish(x), atol=1e-6, rtol=1e-5)
    assert torch.equal(diff.res2img(x.to(DEV), up.to(DEV)).cpu(), O.res2img(x, up, hp))
    want = ((x - up) * hp["res_rescale"]).clamp(-1, 1)
    assert torch.equal(diff.img2res(x.to(DEV), up.to(DEV)).cpu(), want)
    assert torch.equal(diff.res2img(x.to(DEV), up.to(DEV), clip_input=False).cpu(), x / hp["res_rescale"] + up)


def test_p_sample_update_matches_reference_fixture(ours):
    M, diff = ours
    g = golden("p_sample.pt")
    old = M.noise_like
    M.noise_like = lambda shape, device, repeat=False: g["noise"].to(device)
    try:
        xp, x0 = diff.p_sample(g["x"].to(DEV), g["t"].to(DEV), None, None, noise_pred=g["eps"].to(DEV))
    finally:
        M.noise_like = old
    assert torch.equal(x0.cpu(), g["x0"])
    assert torch.allclose(xp.cpu(), g["x_prev"], atol=1e-6, rtol=1e-6)     # expf(0.5*logvar) round-off only
    assert torch.equal(xp.cpu()[0], g["x_prev"][0])                          # t == 0 row: exact, no noise


def test_philox_randn_statistics_and_shard_invariance():
    from dgm_img_super_resolution_b200 import engine as E, _lib as L
    n = 3 * 64 * 64
    a = torch.empty((8, n), device=DEV)
    L.check(L.lib().srd_randn(a.data_ptr(), 8, n, 1234, 0, 0, E._stream()))
    b = torch.empty((4, n), device=DEV)
    L.check(L.lib().srd_randn(b.data_ptr(), 4, n, 1234, 0, 4, E._stream()))      # images 4..7 as a second shard
    assert torch.equal(a[4:], b)
    assert abs(float(a.mean())) < 0.01 and abs(float(a.std()) - 1.0) < 0.01
    assert abs(float((a ** 4).mean()) - 3.0) < 0.1
    c = torch.empty((8, n), device=DEV)
    L.check(L.lib().srd_randn(c.data_ptr(), 8, n, 1234, 1, 0, E._stream()))
    assert abs(float((a * c).mean())) < 0.01                                       # steps are independent


def test_time_table_matches_oracle(ours, cpu_state, hp):
    from dgm_img_super_resolution_b200 import engine as E
    M, diff = ours
    unet = diff.denoise_fn
    t = torch.tensor([0.0, 1.0, 50.0, 99.0, 999.0])
    tab = E.time_table(unet, t).cpu()
    temb = O.time_embedding(cpu_state, t, hp["hidden_size"])
    names = ["downs.%d.%d." % (i, j) for i in range(4) for j in range(2)] + ["mid_block1.", "mid_block2."] + \
            ["ups.%d.%d." % (i, j) for i in range(3) for j in range(2)]
    cols = [F.linear(O.mish(temb), cpu_state["denoise_fn." + n + "mlp.1.weight"],
                     cpu_state["denoise_fn." + n + "mlp.1.bias"]) for n in names]
    want = torch.cat(cols, 1)
    assert tab.shape == want.shape
    assert O.rel_l2(tab, want) < 1e-5


# ------------------------------------------------------------------------------------------ convs
def _mk(M, cls, *a, seed=0, **k):
    torch.manual_seed(seed)
    return cls(*a, **k).to(DEV)


@pytest.mark.parametrize("prec", ["fp32", "bf16_direct", "bf16"])
@pytest.mark.parametrize("cin,cout,H,W,B", [(64, 64, 16, 16, 2), (3, 64, 16, 24, 1), (128, 192, 20, 20, 3),
                                            (256, 256, 10, 12, 2), (64, 128, 34, 18, 1)])
def test_block_reflect_conv_mish(ours, precision, prec, cin, cout, H, W, B):
    M, _ = ours
    precision.precision = prec
    blk = _mk(M, M.Block, cin, cout, groups=0)
    g = torch.Generator().manual_seed(cin + H)
    x = torch.randn((B, cin, H, W), generator=g)
    sd = {k: v.cpu() for k, v in blk.state_dict().items()}
    want = O._block(sd, "", x)
    got = blk(x.to(DEV)).cpu()
    assert O.rel_l2(got, want) < TOL[prec]


@pytest.mark.parametrize("prec", ["fp32", "bf16_direct", "bf16"])
@pytest.mark.parametrize("cin,cout,H,W", [(64, 64, 16, 16), (64, 128, 12, 20), (192, 256, 10, 10)])
def test_resnet_block_without_time(ours, precision, prec, cin, cout, H, W):
    M, _ = ours
    precision.precision = prec
    blk = _mk(M, M.ResnetBlock, cin, cout, time_emb_dim=0, groups=0)
    g = torch.Generator().manual_seed(7)
    x = torch.randn((2, cin, H, W), generator=g)
    sd = {k: v.cpu() for k, v in blk.state_dict().items()}
    h = O._block(sd, "block1.", x)
    h = O._block(sd, "block2.", h)
    want = h + (O._conv(sd, "res_conv", x) if cin != cout else x)
    got = blk(x.to(DEV)).cpu()
    assert O.rel_l2(got, want) < TOL[prec]


@pytest.mark.parametrize("prec", ["fp32", "bf16_direct", "bf16"])
@pytest.mark.parametrize("c,H,W,B", [(64, 16, 16, 2), (128, 20, 12, 2), (192, 40, 40, 2),
                                     # many item rounds with two tiles per item (resident weights), tiles that straddle
                                     # images, ragged column tiles on both sides of the halo box
                                     (64, 96, 104, 5), (256, 12, 20, 3)])
def test_downsample_stride2(ours, precision, prec, c, H, W, B):
    M, _ = ours
    precision.precision = prec
    m = _mk(M, M.Downsample, c)
    g = torch.Generator().manual_seed(8)
    x = torch.randn((B, c, H, W), generator=g)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    want = O._conv(sd, "conv.1", O._rpad(x), stride=2)
    got = m(x.to(DEV)).cpu()
    assert got.shape == want.shape
    assert O.rel_l2(got, want) < TOL[prec]


@pytest.mark.parametrize("prec", ["fp32", "bf16_direct", "bf16"])
@pytest.mark.parametrize("c,H,W", [(64, 8, 8), (128, 10, 6), (192, 20, 20)])
def test_upsample_transposed_conv(ours, precision, prec, c, H, W):
    M, _ = ours
    precision.precision = prec
    m = _mk(M, M.Upsample, c)
    g = torch.Generator().manual_seed(9)
    x = torch.randn((2, c, H, W), generator=g)
    want = F.conv_transpose2d(x, m.conv[0].weight.cpu(), m.conv[0].bias.cpu(), stride=2, padding=1)
    got = m(x.to(DEV)).cpu()
    assert got.shape == want.shape
    assert O.rel_l2(got, want) < TOL[prec]


# ------------------------------------------------------------------------------------------ networks
@pytest.mark.parametrize("prec", ["fp32", "bf16_direct", "bf16"])
@pytest.mark.parametrize("name", ["nets_small.pt", "nets_odd.pt"])
def test_rrdb_net_matches_reference_fixture(ours, precision, prec, name):
    M, diff = ours
    precision.precision = prec
    g = golden(name)
    out, feas = diff.rrdb(g["img_lr"].to(DEV), True)
    assert len(feas) == 9
    tol = 2e-5 if prec == "fp32" else 1e-2        # 126 chained convs: bf16 storage error accumulates (measured 5e-3)
    assert O.rel_l2(torch.stack(feas).cpu(), g["feas"]) < tol
    assert O.rel_l2(out.cpu(), g["rrdb_out"]) < tol
    only = diff.rrdb(g["img_lr"].to(DEV))
    assert torch.equal(only, out)


@pytest.mark.parametrize("prec", ["fp32", "bf16_direct", "bf16"])
@pytest.mark.parametrize("name", ["nets_small.pt", "nets_odd.pt"])
def test_unet_matches_reference_fixture(ours, precision, prec, name):
    M, diff = ours
    precision.precision = prec
    g = golden(name)
    feas = [f.to(DEV) for f in g["feas"]]
    eps = diff.denoise_fn(g["x"].to(DEV), g["t"].to(DEV), feas, g["img_lr_up"].to(DEV))
    assert O.rel_l2(eps.cpu(), g["eps"]) < (1e-4 if prec == "fp32" else 1e-2)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_unet_config1_shape(ours, precision, prec, hp):
    from conftest import gen_inputs
    M, diff = ours
    precision.precision = prec
    g = golden("nets_c1.pt")
    img_lr, img_lr_up, gen = gen_inputs(1, 40, 40, hp["sr_scale"], g["seed_inputs"])
    x = torch.randn((1, 3, 160, 160), generator=gen)
    out, feas = diff.rrdb(img_lr.to(DEV), True)
    eps = diff.denoise_fn(x.to(DEV), g["t"].to(DEV), feas, img_lr_up.to(DEV))
    assert O.rel_l2(eps.cpu(), g["eps"].float()) < (2e-3 if prec == "fp32" else 1e-2)   # fixture is fp16
    assert float(eps.norm()) == pytest.approx(g["eps_norm"], rel=(1e-4 if prec == "fp32" else 1e-2))


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("graph", [False, True])
def test_sampler_T4_matches_reference_fixture(hp, prec, graph):
    """Full `sample()` with injected noise: per-step eps rel-L2 and final-image PSNR (north_star)."""
    from oracle.ref_loader import inject_noise
    M, diff = build_ours(hp, timesteps=4, device=DEV)
    M.runtime.precision, M.runtime.graph, M.runtime.noise = prec, graph, "torch"
    g = golden("sample_T4.pt")
    try:
        with inject_noise(M, g["noises"]):
            img, rrdb_out = diff.sample(g["img_lr"].to(DEV), g["img_lr_up"].to(DEV), tuple(g["img_lr_up"].shape))
        assert O.psnr(img.cpu(), g["img"]) >= (80.0 if prec == "fp32" else 40.0)
        assert O.rel_l2(rrdb_out.cpu(), g["rrdb_out"]) < (2e-5 if prec == "fp32" else 1e-2)
        if not graph:
            with inject_noise(M, g["noises"]):
                img2, _, inter = diff.sample(g["img_lr"].to(DEV), g["img_lr_up"].to(DEV),
                                             tuple(g["img_lr_up"].shape), save_intermediate=True)
            assert len(inter) == 4 and torch.equal(inter[-1][0], img2.cpu())
            assert O.psnr(img2.cpu(), g["img"]) >= (80.0 if prec == "fp32" else 40.0)
    finally:
        M.runtime.precision, M.runtime.graph = "bf16", True


def test_sampler_per_step_eps_against_oracle(hp, cpu_state):
    """Per-step denoiser output along a coupled trajectory (T=4), bf16 vs the CPU oracle."""
    M, diff = build_ours(hp, timesteps=4, device=DEV)
    g = golden("sample_T4.pt")
    sched = O.make_schedule(4, hp)
    x = g["noises"][0].clone()
    with torch.no_grad():
        _, feas = O.rrdb_forward(cpu_state, g["img_lr"], hp)
    feas_d = [f.to(DEV) for f in feas]
    for prec, tol in (("fp32", 1e-4), ("bf16", 1e-2)):
        M.runtime.precision = prec
        for k, i in enumerate(reversed(range(4))):
            t = torch.full((2,), i, dtype=torch.long)
            eps = diff.denoise_fn(x.to(DEV), t.to(DEV), feas_d, None).cpu()
            assert O.rel_l2(eps, g["eps"][k]) < tol, (prec, i)
            x, _ = O.p_sample_update(sched, x, t, g["eps"][k], g["noises"][1 + k])
        x = g["noises"][0].clone()
    M.runtime.precision = "bf16"


def test_philox_sampler_runs_and_is_shard_invariant(hp):
    M, diff = build_ours(hp, timesteps=3, device=DEV)
    M.runtime.noise = "philox"
    try:
        from conftest import gen_inputs
        img_lr, img_lr_up, _ = gen_inputs(4, 8, 8, hp["sr_scale"], 5)
        M.runtime.seed_counter = 41
        full, _ = diff.sample(img_lr.to(DEV), img_lr_up.to(DEV), tuple(img_lr_up.shape))
        M.runtime.seed_counter = 41
        M.runtime.image_offset = 2
        half, _ = diff.sample(img_lr[2:].to(DEV), img_lr_up[2:].to(DEV), tuple(img_lr_up[2:].shape))
        assert torch.isfinite(full).all()
        assert torch.equal(full[2:], half)          # image k of a batch == the same image in a shard
    finally:
        M.runtime.noise, M.runtime.image_offset = "torch", 0


@pytest.mark.parametrize("a_mode,what", [(1 | 64, "ticket-based item scheduling"), (1 | 4, "single-CTA MMAs"),
                                         (1 | 16, "resident weights with MT = 1"), (1 | 8, "no dependent launch"),
                                         (1 | 2, "register epilogue"), (0, "per-kx activation boxes")])
def test_tc_mode_switches_do_not_change_results(ours, a_mode, what):
    """The tuning / validation switches of srd_tc_set_mode change HOW the tcgen05 kernel works, never what it computes:
    a ResnetBlock with enough tiles for several item rounds per CTA pair gives the same bits in every scheduling /
    pairing / launch mode (64 -> 128 channels: 1x1 res_conv as a second output), and the same values up to the
    rounding order of the residual add with the register epilogue or the per-kx activation boxes (64 -> 64: those two
    modes do not offer the second output)."""
    from dgm_img_super_resolution_b200 import _lib as L
    M, _ = ours
    M.runtime.precision = "bf16"
    exact = a_mode not in (1 | 2, 0)
    g = torch.Generator().manual_seed(97)
    x = torch.randn((48, 64, 40, 40), generator=g).to(DEV)
    outs = []
    try:
        for mode in (1, a_mode):
            L.check(L.lib().srd_tc_set_mode(mode, 1), "srd_tc_set_mode")
            blk = _mk(M, M.ResnetBlock, 64, 128 if exact else 64, time_emb_dim=0, groups=0, seed=5)   # new plans
            outs.append(blk(x))
    finally:
        L.check(L.lib().srd_tc_set_mode(1, 1), "srd_tc_set_mode")
    if exact:
        assert torch.equal(outs[0], outs[1]), what
    else:
        assert O.rel_l2(outs[1].cpu(), outs[0].cpu()) < 3e-3, what
    assert L.lib().srd_tc_set_mode(128, 1) != 0 and L.lib().srd_watchdog_flag() == 0


def test_sample_sharded_is_shard_invariant_with_real_diffusion(hp):
    """dist.sample_sharded with a REAL GaussianDiffusion: (a) through an in-process world-size-1 group it equals a
    plain Philox `sample()` with the job seed; (b) the two halves of the batch, sampled separately the way two ranks
    would (slice start as image offset, same job seed), concatenate to exactly the single-GPU result."""
    import os
    import torch.distributed as dist
    from dgm_img_super_resolution_b200 import dist as D
    M, diff = build_ours(hp, timesteps=3, device=DEV)
    img_lr, img_lr_up, _ = gen_inputs(5, 8, 8, hp["sr_scale"], 33)
    lr, up = img_lr.to(DEV), img_lr_up.to(DEV)
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29571")
        dist.init_process_group("gloo", rank=0, world_size=1)
        created = True
    try:
        M.runtime.noise = "torch"                      # sample_sharded must override this
        M.runtime.seed_override = 424242
        whole, whole_rrdb = D.sample_sharded(diff, lr, up)
        assert M.runtime.noise == "torch" and M.runtime.seed_override == 424242      # restored
        M.runtime.noise = "philox"
        ref, _ = diff.sample(lr, up, tuple(up.shape))
        assert torch.equal(whole, ref)
        parts = []
        for rank in range(2):
            start, count = D.shard_range(5, 2, rank)
            M.runtime.image_offset = start
            img, _ = diff.sample(lr[start:start + count], up[start:start + count], (count, 3, 32, 32))
            parts.append(img)
        assert torch.equal(torch.cat(parts, 0), whole)
    finally:
        M.runtime.noise, M.runtime.seed_override, M.runtime.image_offset = "torch", None, 0
        if created:
            dist.destroy_process_group()


def test_n_split_channel_groups_match_oracle(hp):
    """SrdConvDesc.n_split = 2 (opt-in, SRD_NSPLIT=1): a 256-channel Block computed as two 128-channel groups that
    share every weight block between two tile pairs -- against the oracle, and bit-equal to the unsplit launch."""
    from dgm_img_super_resolution_b200 import engine as E
    import dgm_img_super_resolution_b200.model as M
    M.hparams.clear()
    M.hparams.update(hp)
    torch.manual_seed(7)
    blk = M.Block(192, 256, groups=0).to(DEV)
    x = torch.randn((3, 192, 20, 12), generator=torch.Generator().manual_seed(8))
    want = O._block({"block.1.weight": blk.block[1].weight.detach().cpu(), "block.1.bias": blk.block[1].bias.detach().cpu()},
                    "", x)
    M.runtime.precision = "bf16"
    old = E.NSPLIT
    try:
        E.NSPLIT = False
        plain = blk(x.to(DEV)).cpu()
        E.NSPLIT = True
        split = blk(x.to(DEV)).cpu()
    finally:
        E.NSPLIT = old
    assert O.rel_l2(split, want) < 1e-2
    assert torch.equal(split, plain)
