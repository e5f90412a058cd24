"""Stand-alone module forwards and sampler helper methods (row P / R symbols a drop-in caller may touch:
`/root/reference/srdiff/model.py:398-434, 670-706, 962-990`).  Fixture: tests/golden/api_rows.pt, generated from
the unmodified reference by oracle/make_golden_api.py.  CPU tests pin the oracle's arithmetic to it; GPU tests
compare the CUDA path (public API -> C ABI) with it."""
import copy

import pytest
import torch

from conftest import build_ours, golden
from oracle import srdiff_oracle as O

DEV = "cuda:0"


# ------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_p_sample_pieces_match_helper_fixture(hp):
    h = golden("api_rows.pt")["helpers"]
    s = O.make_schedule(100, hp)
    _, x0 = O.p_sample_update(s, h["x"], h["t"], h["eps"], torch.zeros_like(h["x"]))
    assert torch.equal(x0, h["pmv_x0"])
    mean, _ = O.p_sample_update(s, h["x"], h["t"], h["eps"], torch.zeros_like(h["x"]))
    assert torch.equal(mean, h["pmv_mean"])


def test_oracle_rdb_matches_fixture(hp, cpu_state):
    r = golden("api_rows.pt")["rdb"]
    with torch.no_grad():
        out = O._rdb5c(cpu_state, "rrdb.RRDB_trunk.0.RDB1.", r["x"])
    assert O.rel_l2(out, r["rdb1"]) < 1e-6


def test_inference_only_module_raises_and_copies(hp):
    """GaussianDiffusion(...) is not callable as a training step (explicit error, not nn.Module's default), and a
    model stays deep-copyable / picklable (EMA copies, torch.save(model)) -- the engine caches are dropped."""
    M, diff = build_ours(hp, timesteps=4)
    with pytest.raises(NotImplementedError):
        diff(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 2, 2), torch.zeros(1, 3, 8, 8))
    diff._plans["x"] = object()
    diff.denoise_fn._progs["p"] = (1, 2)
    diff.denoise_fn.cond_proj.__dict__["_srd_pack"] = object()
    c = copy.deepcopy(diff)
    assert c._plans == {} and c.denoise_fn._progs == {} and "_srd_pack" not in c.denoise_fn.cond_proj.__dict__
    assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), diff.state_dict().values()))
    import io
    import pickle
    buf = io.BytesIO()
    pickle.dump(diff, buf)


def test_timestep_out_of_range_raises(hp):
    M, diff = build_ours(hp, timesteps=4)
    x = torch.zeros(2, 3, 4, 4)
    with pytest.raises(Exception):
        diff._check_t(torch.tensor([0, 4]), "t")
    with pytest.raises(Exception):
        diff._check_t(torch.tensor([-1, 1]), "t")
    diff._check_t(torch.tensor([-1, 3]), "t", allow_negative=True)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_sampler_helpers_match_reference_fixture(hp):
    h = golden("api_rows.pt")["helpers"]
    M, diff = build_ours(hp, device=DEV)
    x, e, t = h["x"].to(DEV), h["eps"].to(DEV), h["t"].to(DEV)
    mean, var, logvar = diff.q_mean_variance(x, t)
    assert torch.equal(mean.cpu(), h["q_mean"]) and torch.equal(var.cpu(), h["q_var"])
    assert torch.equal(logvar.cpu(), h["q_logvar"])
    assert torch.equal(diff.predict_start_from_noise(x, t, e).cpu(), h["x0"])
    pm, pv, plv = diff.q_posterior(e, x, t)
    assert torch.equal(pm.cpu(), h["post_mean"]) and torch.equal(pv.cpu(), h["post_var"])
    assert torch.equal(plv.cpu(), h["post_logvar"])
    mm, mv, mlv, xr = diff.p_mean_variance(x, t, e, True)
    assert torch.equal(mm.cpu(), h["pmv_mean"]) and torch.equal(xr.cpu(), h["pmv_x0"])
    assert torch.equal(mv.cpu(), h["pmv_var"]) and torch.equal(mlv.cpu(), h["pmv_logvar"])
    mm2, _, _, xr2 = diff.p_mean_variance(x, t, e, False)
    assert torch.equal(mm2.cpu(), h["pmv_mean_noclip"]) and torch.equal(xr2.cpu(), h["pmv_x0_noclip"])
    with pytest.raises(Exception):
        diff.p_sample(x, torch.full((4,), 100, device=DEV), None, None, noise_pred=e)
    with pytest.raises(Exception):
        diff.q_sample(x, torch.full((4,), 100, device=DEV))


@pytest.mark.gpu
def test_pos_emb_and_wrappers_match_reference_fixture(hp):
    g = golden("api_rows.pt")
    M, diff = build_ours(hp, device=DEV)
    pe = g["pos_emb"]
    out = M.SinusoidalPosEmb(pe["dim"])(pe["t"].to(DEV)).cpu()
    assert out.shape == pe["out"].shape and float((out - pe["out"]).abs().max()) < 2e-6
    w = g["wrappers"]
    x = w["x"].to(DEV)
    assert float((M.Residual(M.Mish())(x).cpu() - w["residual_mish"]).abs().max()) < 1e-6
    rz = M.Rezero(M.Mish()).to(DEV)
    rz.g.data.fill_(0.37)
    assert float((rz(x).cpu() - w["rezero_mish_g037"]).abs().max()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("bf16", 1e-2), ("bf16_direct", 1e-2)])
def test_dense_blocks_standalone_forward(hp, prec, tol):
    r = golden("api_rows.pt")["rdb"]
    M, diff = build_ours(hp, device=DEV)
    M.runtime.precision = prec
    try:
        blk = diff.rrdb.RRDB_trunk[0]
        x = r["x"].to(DEV)
        assert O.rel_l2(blk.RDB1(x).cpu(), r["rdb1"]) < tol
        assert O.rel_l2(blk(x).cpu(), r["rrdb"]) < tol
    finally:
        M.runtime.precision = "bf16"
