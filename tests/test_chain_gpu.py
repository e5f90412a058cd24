"""Layer chain and fused segments (include/srdiff_b200.h: srd_chain_*; both opt-in, SRD_CHAIN=1 / SRD_FUSE=1): the whole Unet.forward as ONE persistent
launch whose layers are ordered by completion counters instead of kernel boundaries.  It must be bit-identical to the
same plans launched one by one -- same tiles, same arithmetic -- on every run (a missed dependency shows up as a
mismatch or as run-to-run differences), at small / ragged shapes and at the bench shape."""
import pytest
import torch

from conftest import build_ours, gen_inputs
from dgm_img_super_resolution_b200 import engine as E

DEV = "cuda:0"
pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _chain_on():
    old = E.USE_CHAIN
    E.USE_CHAIN = True
    yield
    E.USE_CHAIN = old


def _plan(hp, B, h, w, seed, timesteps=4):
    M, diff = build_ours(hp, timesteps=timesteps, device=DEV)
    M.runtime.precision = "bf16"
    img_lr, img_lr_up, g = gen_inputs(B, h, w, hp["sr_scale"], seed)
    plan = diff._plan(B, h, w, torch.device(DEV))
    plan.img_lr.copy_(img_lr.to(DEV))
    plan.img_lr_up.copy_(img_lr_up.to(DEV))
    plan.x.copy_(torch.randn(plan.x.shape, generator=g).to(DEV))
    plan.encode()
    plan.load_x()
    torch.cuda.synchronize()
    return M, diff, plan


def _run(plan, t, chain):
    bld = plan.unet_prog.bld
    saved = bld.chain
    if not chain:
        bld.chain = None
    try:
        plan.unet_prog.eps.fill_(float("nan"))
        plan.unet_prog.run(plan.temb.data_ptr() + 4 * t * plan.temb.shape[1])
        torch.cuda.synchronize()
        return plan.unet_prog.eps.clone()
    finally:
        bld.chain = saved


@pytest.mark.parametrize("B,h,w", [(2, 8, 8), (3, 6, 10), (1, 4, 4), (5, 16, 12)])
def test_chain_equals_layer_by_layer_small(hp, B, h, w):
    M, diff, plan = _plan(hp, B, h, w, 41)
    assert plan.unet_prog.bld.chain is not None, "the default bf16 Unet program must run as a chain"
    assert plan.unet_prog.bld.chain.n_segments < len(plan.unet_prog.bld.ops)
    for t in (3, 0):
        ref = _run(plan, t, chain=False)
        assert torch.isfinite(ref).all()
        for _ in range(3):
            assert torch.equal(_run(plan, t, chain=True), ref)


def test_chain_single_layer_segments_equal(hp):
    old = E.CHAIN_SINGLE_SEGMENTS
    E.CHAIN_SINGLE_SEGMENTS = True
    try:
        M, diff, plan = _plan(hp, 2, 8, 8, 43)
        bld = plan.unet_prog.bld
        assert bld.chain.n_segments == len(bld.ops)
        assert torch.equal(_run(plan, 2, chain=True), _run(plan, 2, chain=False))
    finally:
        E.CHAIN_SINGLE_SEGMENTS = old


def test_chain_equals_layer_by_layer_bench_shape(hp):
    """configs[1]: batch 64 of 160x160 -- every layer spans many item rounds, all 148 SMs carry several layers at
    once; five runs must agree bit for bit with the layer-by-layer launches."""
    M, diff, plan = _plan(hp, 64, 40, 40, 45, timesteps=100)
    ref = _run(plan, 57, chain=False)
    for _ in range(5):
        assert torch.equal(_run(plan, 57, chain=True), ref)


def test_chain_inside_the_sampler_graph(hp):
    """sample() through the CUDA graph with the chain == sample() with per-layer launches (same Philox seed)."""
    M, diff = build_ours(hp, timesteps=5, device=DEV)
    M.runtime.noise, M.runtime.precision, M.runtime.graph = "philox", "bf16", True
    img_lr, img_lr_up, _ = gen_inputs(4, 12, 12, hp["sr_scale"], 47)
    try:
        M.runtime.seed_override = 777
        a, _ = diff.sample(img_lr.to(DEV), img_lr_up.to(DEV), tuple(img_lr_up.shape))
        assert diff.last_stats["gpu_launches"] < 40 * 5            # one launch per denoiser call, not ~42
        old = E.USE_CHAIN
        E.USE_CHAIN = False
        try:
            M2, diff2 = build_ours(hp, timesteps=5, device=DEV)
            b, _ = diff2.sample(img_lr.to(DEV), img_lr_up.to(DEV), tuple(img_lr_up.shape))
        finally:
            E.USE_CHAIN = old
        assert torch.equal(a, b)
    finally:
        M.runtime.noise, M.runtime.seed_override = "torch", None


# ------------------------------------------------------------------------------------------ fused segments (default)
def _run_exec(plan, t, fused):
    bld = plan.unet_prog.bld
    saved = bld.exec
    if not fused:
        bld.exec = None
    try:
        plan.unet_prog.eps.fill_(float("nan"))
        plan.unet_prog.run(plan.temb.data_ptr() + 4 * t * plan.temb.shape[1])
        torch.cuda.synchronize()
        return plan.unet_prog.eps.clone()
    finally:
        bld.exec = saved


@pytest.mark.parametrize("B,h,w", [(2, 8, 8), (3, 6, 10), (5, 16, 12), (64, 40, 40)])
def test_fused_segments_equal_layer_by_layer(hp, B, h, w):
    """The engine's default: runs of same-plan layers as fused-segment launches (compile-time unrolled layer loop,
    epoch-scaled completion counters that are never cleared).  Bit-identical to one launch per layer, on every one of
    several consecutive launches (each advances the device-side epoch), also at the bench shape."""
    E.USE_CHAIN = False
    old_fuse, E.FUSE_SEGMENTS = E.FUSE_SEGMENTS, True
    try:
        M, diff, plan = _plan(hp, B, h, w, 51, timesteps=100 if B == 64 else 4)
    finally:
        E.FUSE_SEGMENTS = old_fuse
    bld = plan.unet_prog.bld
    assert bld.chain is None and bld.exec is not None
    fused = [op for op in bld.exec if isinstance(op, E.ChainOp)]
    assert fused and all(op.fused and 2 <= len(op.ops) <= 3 for op in fused)       # K = 2 (default build) or 3
    assert sum(len(op.ops) if isinstance(op, E.ChainOp) else 1 for op in bld.exec) == len(bld.ops)
    ref = _run_exec(plan, 2, fused=False)
    assert torch.isfinite(ref).all()
    for _ in range(4):
        assert torch.equal(_run_exec(plan, 2, fused=True), ref)
    if B == 64:
        assert len(bld.exec) <= 32, len(bld.exec)          # 42 layers -> 27 launches with K = 3, 30 with K = 2
