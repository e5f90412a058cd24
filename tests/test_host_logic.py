"""CPU tests of the host side: weight packing / tap tables (via tests/emulate.py), state-dict
contract, schedule buffers, C-ABI symbol export, loud failure without CUDA."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

import emulate as EM
from conftest import ROOT, build_ours, golden


def _nchw(o):
    return o.permute(0, 3, 1, 2)


@pytest.mark.parametrize("cins,O,k,stride,reflect", [
    ([64], 64, 3, 1, True), ([3], 64, 3, 1, True), ([96, 64], 48, 3, 1, True), ([32, 16, 16], 16, 3, 1, False),
    ([64], 64, 3, 2, True), ([130], 24, 3, 2, True), ([64], 128, 1, 1, True), ([100, 92], 3, 1, 1, False)])
def test_pack_conv_layouts(cins, O, k, stride, reflect):
    from dgm_img_super_resolution_b200 import engine as E
    torch.manual_seed(0)
    I = sum(cins)
    x = torch.randn(2, I, 8, 12)
    w = torch.randn(O, I, k, k) * 0.1
    if k == 3:
        xp = F.pad(x, (1, 1, 1, 1), mode="reflect" if reflect else "constant")
        want = F.conv2d(xp, w, stride=stride)
    else:
        want = F.conv2d(x, w)
    wd, wt = E._pack_conv(w, cins, stride)
    srcs, a = [], 0
    for c in cins:
        srcs.append(EM.to_padded_nhwc(x[:, a:a + c], reflect))
        a += c
    k0 = (3 - k) // 2
    got_d = _nchw(EM.emulate_direct(srcs, cins, wd, O, stride, k, k0, 1))
    assert torch.allclose(got_d, want, atol=1e-4, rtol=1e-4)
    got_t = _nchw(EM.emulate_tc(srcs, cins, wt, O, stride, k, k0, 1))
    assert torch.allclose(got_t, want, atol=0.08, rtol=0.02)        # weights rounded to bf16
    assert wt.shape[-2] % 16 == 0 and wt.shape[-1] == 64 and wt.dtype == torch.bfloat16


def test_pack_conv_dual_interleaves_the_1x1_blocks():
    """SrdConvDesc.N2: per (source, chunk) nine 3x3 k-blocks then the 1x1 block of the same channels."""
    from dgm_img_super_resolution_b200 import engine as E
    torch.manual_seed(2)
    cins = [100, 64]
    w3, w1 = torch.randn(64, 164, 3, 3), torch.randn(64, 164, 1, 1)
    wt = E._pack_conv_dual(w3, w1, cins).float()
    _, ref3 = E._pack_conv(w3, cins, 1)
    _, ref1 = E._pack_conv(w1, cins, 1)
    assert wt.shape == (3 * 10, 64, 64)
    for ch in range(3):
        assert torch.equal(wt[ch * 10:ch * 10 + 9], ref3.float()[ch * 9:ch * 9 + 9])
        assert torch.equal(wt[ch * 10 + 9], ref1.float()[ch])


def test_pack_conv_win8_windows():
    """SRD_SRC_WIN8: the first Unet conv as three K=64 blocks over 8-pixel windows of the 8-channel buffer
    (K index = dx*8 + c; the window of the last pixels of a row runs into the next row on zero weights)."""
    from dgm_img_super_resolution_b200 import engine as E
    torch.manual_seed(1)
    x = torch.randn(2, 3, 8, 16)
    w = torch.randn(64, 3, 3, 3) * 0.2
    want = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)
    wt = E._pack_conv_win8(w).float()                                   # [3 ky][Npad][64]
    assert wt.shape == (3, 64, 64)
    xp = EM.to_padded_nhwc(x, True)                                     # [B][H+2][W+2][3] -> pad channels to 8
    B, Hp, Wp, _ = xp.shape
    buf = torch.zeros(B, Hp, Wp, 8)
    buf[..., :3] = xp
    flat = torch.cat([buf.reshape(-1), torch.zeros(64)])                # + the slack the engine allocates
    out = torch.zeros(B, 64, 8, 16)
    for b in range(B):
        for y in range(8):
            for xx in range(16):
                for ky in range(3):
                    o = ((b * Hp + y + ky) * Wp + xx) * 8               # window starts at the tap's column
                    out[b, :, y, xx] += wt[ky] @ flat[o:o + 64]
    assert torch.allclose(out, want, atol=0.08, rtol=0.02)              # weights rounded to bf16


@pytest.mark.parametrize("cins,O,S", [([64], 64, 2), ([192], 40, 2), ([32, 32, 32], 64, 4), ([70, 26], 16, 4)])
def test_pack_conv_transpose_phases(cins, O, S):
    from dgm_img_super_resolution_b200 import engine as E
    torch.manual_seed(1)
    I = sum(cins)
    x = torch.randn(2, I, 5, 7)
    w = torch.randn(I, O, 2 * S, 2 * S) * 0.1
    want = F.conv_transpose2d(x, w, stride=S, padding=S // 2)
    wd, wt = E._pack_convT(w, cins, S)
    srcs, a = [], 0
    for c in cins:
        srcs.append(EM.to_padded_nhwc(x[:, a:a + c], False))
        a += c
    got_d = _nchw(EM.emulate_direct(srcs, cins, wd, O, 1, 2, 0, S))
    assert got_d.shape == want.shape
    assert torch.allclose(got_d, want, atol=1e-4, rtol=1e-4)
    got_t = _nchw(EM.emulate_tc(srcs, cins, wt, O, 1, 2, 0, S))
    assert torch.allclose(got_t, want, atol=0.08, rtol=0.02)


def test_state_dict_contract(hp):
    M, diff = build_ours(hp)
    sd = diff.state_dict()
    chk = __import__("json").load(open(os.path.join(ROOT, "tests", "golden", "state_checksum.json")))
    assert set(sd) == set(chk) and len(sd) == 396
    for k, v in sd.items():
        assert float(v.double().sum()) == pytest.approx(chk[k][0], rel=1e-9, abs=1e-9), k
        assert float(v.double().abs().sum()) == pytest.approx(chk[k][1], rel=1e-9), k
    g = golden("schedules.pt")["cosine_T100"]
    for k, v in g.items():
        assert torch.equal(sd[k], v), k
    # load_state_dict round trip, strict and non-strict (st_app1.py:1645-1648)
    M2, diff2 = build_ours(hp, seed=99)
    diff2.load_state_dict(sd, strict=True)
    diff2.load_state_dict({k: v for k, v in sd.items() if "rrdb" in k}, strict=False)
    assert diff.num_timesteps == 100 and diff.sample_tqdm is False and hasattr(diff, "denoise_fn") and hasattr(diff, "rrdb")


def test_other_timesteps_and_schedules(hp):
    g = golden("schedules.pt")
    M, d25 = build_ours(hp, timesteps=25)
    assert all(torch.equal(getattr(d25, k), v) for k, v in g["cosine_T25"].items())
    M, dl = build_ours(dict(hp, beta_schedule="linear"), timesteps=100)
    assert all(torch.equal(getattr(dl, k), v) for k, v in g["linear_T100"].items())
    with pytest.raises(NotImplementedError):
        M.get_beta_schedule(10, "nope")


def test_unsupported_config_is_loud(hp):
    with pytest.raises(KeyError):
        import dgm_img_super_resolution_b200.model as M
        M.hparams.clear()
        M.RRDBNet(3, 3, 32, 8, 16)          # reference: KeyError on missing hparams (:720)


def test_library_exports_every_declared_symbol():
    from dgm_img_super_resolution_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "libsrdiff_b200.so is not built: run __graft_entry__.build()"
    hdr = open(os.path.join(ROOT, "include", "srdiff_b200.h")).read()
    declared = set(re.findall(r"\b(srd_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    L = _lib.lib()
    assert L.srd_version() >= 100
    # struct sizes must agree with the header (x86-64 ABI): spot-check through an invalid call
    d = _lib.SrdConvDesc()
    h = ctypes.c_void_p()
    assert L.srd_conv_plan_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b"conv" in L.srd_last_error()


def test_no_cpu_fallback(hp):
    M, diff = build_ours(hp)
    from dgm_img_super_resolution_b200._lib import SrdError
    x = torch.zeros(1, 3, 8, 8)
    with pytest.raises(SrdError):
        diff.rrdb(x)
    with pytest.raises(SrdError):
        diff.sample(x, torch.zeros(1, 3, 32, 32), (1, 3, 32, 32))
    with pytest.raises(SrdError):
        M.Mish()(x)


def test_committed_profile_evidence_is_consistent(tmp_path):
    """The per-pass traffic file bench.py reports as roofline.traffic is what tools/ncu_summary.py extracts from the
    committed ncu launch list (42 conv launches between two p_sample launches)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rnd in ("r01", "r02"):
        csv_path = os.path.join(root, "profiles", rnd, "ncu_launch_list_bench_warmL2.csv")
        out = str(tmp_path / ("pass_%s.json" % rnd))
        subprocess.check_call([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), csv_path, "--pass-json", out],
                              stdout=subprocess.DEVNULL)
        got = json.load(open(out))
        committed = json.load(open(os.path.join(root, "profiles", rnd, "unet_pass_traffic.json")))
        assert got["conv_launches_per_unet_pass"] == committed["conv_launches_per_unet_pass"] == 42
        for k in ("dram_bytes_read", "dram_bytes_write", "serialized_cold_us"):
            assert got[k] == pytest.approx(committed[k], rel=1e-9)
        assert 4.5e9 < committed["dram_bytes_read"] + committed["dram_bytes_write"] < 6.5e9
