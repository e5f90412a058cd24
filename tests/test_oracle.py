"""CPU tests: the oracle restatement against the golden vectors generated from the live
reference (oracle/make_golden.py), and against the reference itself when it is present."""
import pytest
import torch

from conftest import golden, gen_inputs
from oracle import ref_loader
from oracle import srdiff_oracle as O

TOL = 2e-6   # fp32 round-off between two orderings of the same ATen ops


def test_schedules_match_reference_fixtures(hp):
    g = golden("schedules.pt")
    for name, T, over in (("cosine_T100", 100, {}), ("cosine_T25", 25, {}), ("cosine_T1000", 1000, {}),
                          ("linear_T100", 100, {"beta_schedule": "linear"})):
        s = O.make_schedule(T, dict(hp, **over))
        assert set(s) == set(g[name])
        for k in s:
            assert torch.equal(s[k], g[name][k]), (name, k)
    # spot values quoted in SURVEY 8a row G
    s = O.make_schedule(100, hp)
    assert abs(float(s["betas"][0]) - 6.3128e-4) < 1e-7 and float(s["betas"][99]) == pytest.approx(0.999)
    assert float(s["sqrt_recip_alphas_cumprod"][99]) == pytest.approx(2029.2, rel=1e-4)
    assert float(s["posterior_mean_coef1"][0]) == 1.0 and float(s["posterior_mean_coef2"][0]) == 0.0


@pytest.mark.parametrize("name", ["nets_small.pt", "nets_odd.pt"])
def test_rrdb_and_unet_match_fixture(hp, cpu_state, name):
    g = golden(name)
    with torch.no_grad():
        out, feas = O.rrdb_forward(cpu_state, g["img_lr"], hp)
        eps = O.unet_forward(cpu_state, g["x"], g["t"], feas, hp)
    assert O.rel_l2(out, g["rrdb_out"]) < TOL
    assert O.rel_l2(torch.stack(feas), g["feas"]) < TOL
    assert O.rel_l2(eps, g["eps"]) < 5e-6


def test_c1_shape_matches_fixture(hp, cpu_state):
    g = golden("nets_c1.pt")
    img_lr, img_lr_up, gen = gen_inputs(1, 40, 40, hp["sr_scale"], g["seed_inputs"])
    x = torch.randn((1, 3, 160, 160), generator=gen)
    with torch.no_grad():
        out, feas = O.rrdb_forward(cpu_state, img_lr, hp)
        eps = O.unet_forward(cpu_state, x, g["t"], feas, hp)
    assert float(eps.norm()) == pytest.approx(g["eps_norm"], rel=1e-5)
    assert O.rel_l2(eps, g["eps"].float()) < 1e-3        # fixture stored as fp16
    assert O.rel_l2(out, g["rrdb_out"].float()) < 1e-3


def test_p_sample_update_matches_fixture(hp):
    g = golden("p_sample.pt")
    s = O.make_schedule(100, hp)
    xp, x0 = O.p_sample_update(s, g["x"], g["t"], g["eps"], g["noise"])
    assert torch.equal(xp, g["x_prev"]) and torch.equal(x0, g["x0"])
    assert torch.equal(O.res2img(xp, g["x"], hp), g["res2img"])
    # t == 0 rows get no noise
    xp2, _ = O.p_sample_update(s, g["x"], g["t"], g["eps"], g["noise"] * 3)
    assert torch.equal(xp2[0], xp[0]) and not torch.equal(xp2[1], xp[1])


def test_sampler_matches_fixture(hp, cpu_state):
    g = golden("sample_T4.pt")
    s = O.make_schedule(4, hp)
    rec = {}
    with torch.no_grad():
        img, rrdb_out = O.sample(cpu_state, s, hp, g["img_lr"], g["img_lr_up"], g["noises"], record=rec)
    assert O.rel_l2(torch.stack(rec["eps"]), g["eps"]) < 1e-5
    assert O.rel_l2(rrdb_out, g["rrdb_out"]) < TOL
    assert O.psnr(img, g["img"]) > 100.0


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
def test_oracle_against_live_reference(hp, cpu_state):
    ref, diff = ref_loader.build_reference(ref_loader.load_hparams(), seed=1234)
    rsd = diff.state_dict()
    assert all(torch.equal(rsd[k], cpu_state[k]) for k in rsd)
    img_lr, img_lr_up, gen = gen_inputs(2, 12, 8, hp["sr_scale"], 23)
    x = torch.randn((2, 3, 48, 32), generator=gen)
    t = torch.tensor([98, 1])
    with torch.no_grad():
        r_out, r_feas = diff.rrdb(img_lr, True)
        r_eps = diff.denoise_fn(x, t, r_feas, img_lr_up)
        o_out, o_feas = O.rrdb_forward(cpu_state, img_lr, hp)
        o_eps = O.unet_forward(cpu_state, x, t, o_feas, hp)
    assert O.rel_l2(o_out, r_out) < TOL and O.rel_l2(o_eps, r_eps) < 5e-6


def test_oracle_full_T100_trajectory_matches_reference_fixture(hp, cpu_state):
    """configs[0] end to end: one 40x40 -> 160x160 image through all 100 steps (fixture: final image of the
    unmodified reference, fp16; inputs and the 101 noise tensors are regenerated from the stored seeds)."""
    from oracle.make_golden_T100 import make_noises
    f = golden("sample_T100.pt")
    img_lr, img_lr_up, _ = gen_inputs(1, 40, 40, hp["sr_scale"], f["seed_inputs"])
    noises = make_noises(f["T"], img_lr_up.shape)
    with torch.no_grad():
        img, _ = O.sample(cpu_state, O.make_schedule(f["T"], hp), hp, img_lr, img_lr_up, noises)
    assert float(img.norm()) == pytest.approx(f["img_norm"], rel=1e-4)
    assert O.psnr(img, f["img"].float()) > 100.0
