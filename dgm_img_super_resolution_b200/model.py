"""Drop-in mirror of the reference's `srdiff/model.py` module API for the reverse-diffusion
inference path: the module-global ``hparams`` dict, ``RRDBNet`` / ``Unet`` /
``GaussianDiffusion`` with the reference's constructor signatures, parameter names and shapes
(state-dict compatible, Appendix C), and ``sample / p_sample / res2img / img2res``.

Parameters live in ordinary nn.Conv2d / nn.Linear containers (fp32, NCHW) -- that is the API
truth -- but no forward ever calls ATen arithmetic: every forward converts to padded NHWC and
launches the sm_100a kernels of libsrdiff_b200.so through `engine.py`.  CPU tensors raise.

Line numbers cite /root/reference/srdiff/model.py.
"""
import math
from functools import partial

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import engine as E

# module-global config dict, filled by the caller from srdiff/config.yaml (:274; st_app1.py:1568-1578)
hparams = {}


class Runtime:
    """Knobs that do not exist in the reference (it has one precision and one device path)."""
    precision = "bf16"        # "bf16" (tcgen05 path) | "fp32" (validation mode, CUDA-core fp32) | "bf16_direct"
    noise = "torch"           # "torch": x_T / per-step noise via torch.randn + noise_like (reference RNG order)
    #                           "philox": generated inside the p_sample kernel (throughput mode)
    graph = True              # capture encode + T-step loop in one CUDA graph
    noise_budget_bytes = 8 << 30
    plan_cache_size = 4       # SamplerPlans kept per GaussianDiffusion (LRU; one per (B, h, w, precision, device))
    graph_cache_size = 8      # captured CUDA graphs kept per SamplerPlan (LRU; one per noise mode / image offset)
    seed_counter = 0
    image_offset = 0          # global index of image 0 of this shard (multi-GPU shard invariance)
    seed_override = None      # Philox seed forced for the next sample() calls (dist.sample_sharded: one job-wide seed)


runtime = Runtime()


def next_seed(device):
    """Philox seed of the next sampler call: derived from the device's torch generator seed (so torch.manual_seed
    steers it) and a per-process call counter; `runtime.seed_override` pins it instead."""
    if runtime.seed_override is not None:
        return int(runtime.seed_override) & 0x7FFFFFFFFFFFFFFF
    runtime.seed_counter += 1
    if getattr(device, "type", "cuda") != "cuda":          # host-logic tests (gloo): no CUDA generator to read
        base = torch.initial_seed()
    else:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        base = torch.cuda.default_generators[idx].initial_seed()
    return (base * 1000003 + runtime.seed_counter) & 0x7FFFFFFFFFFFFFFF

_CACHE_ATTRS = ("_srd_pack", "_srd_pack_nobias", "_srd_wn", "_srd_zero_bias", "_srd_zero_bias2", "_srd_neg",
                "_last_table")


class _EngineCached(nn.Module):
    """Mixin for the modules that own engine state (plan / program caches holding ctypes handles and CUDA graphs):
    `copy.deepcopy(model)` (EMA copies) and pickling (`torch.save(model)`) carry the parameters only -- the copy
    re-plans on first use."""
    _cache_fields = ()

    def _strip(self, state):
        state = dict(state)
        for k in self._cache_fields:
            if k in state:
                state[k] = {}
        for k in _CACHE_ATTRS:
            state.pop(k, None)
        return state

    def __getstate__(self):
        return self._strip(self.__dict__)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self._strip(self.__dict__).items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        for m in new.modules():                       # packed-weight caches hung on plain nn.Conv2d containers
            for k in _CACHE_ATTRS:
                m.__dict__.pop(k, None)
        return new


def initialize_weights(net_l, scale=0.1):
    """:78-95 -- parameter initialisation only (host side): fan-in Kaiming normal scaled by `scale` with zero biases
    for Conv2d / Linear, unit weight and zero bias for BatchNorm2d."""
    for net in (net_l if isinstance(net_l, list) else [net_l]):
        for m in net.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in')
                m.weight.data *= scale
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias.data, 0.0)


def exists(x):
    return x is not None


def default(val, d):
    from inspect import isfunction
    if exists(val):
        return val
    return d() if isfunction(d) else d


# ------------------------------------------------------------------------------------------ blocks
class Mish(nn.Module):
    """:422-424"""

    def forward(self, x):
        E._require_cuda(x, "Mish input")
        xc = x.contiguous().float()
        y = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            L.check(L.lib().srd_mish_f32(xc.data_ptr(), y.data_ptr(), xc.numel(), E._stream()), "srd_mish_f32")
        return y


class SinusoidalPosEmb(nn.Module):
    """:407-419.  Parameter-free; evaluated inside srd_time_table on the hot path."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        """x [B] (any dtype) -> [B, dim] = cat(sin(x * f), cos(x * f)); the frequencies are evaluated on the host
        exactly as :415-416 does (fp32 exp of an arange), the sin / cos by srd_sinusoidal_emb."""
        E._require_cuda(x, "SinusoidalPosEmb input")
        with torch.cuda.device(x.device):
            t = x.reshape(-1).float().contiguous()
            half = self.dim // 2
            freqs = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).float().to(x.device)
            out = torch.empty((t.numel(), 2 * half), dtype=torch.float32, device=x.device)
            L.check(L.lib().srd_sinusoidal_emb(t.data_ptr(), t.numel(), self.dim, freqs.data_ptr(), out.data_ptr(),
                                               E._stream()), "srd_sinusoidal_emb")
            return out


def _axpby(x, y, a, b):
    """a*x + b*y on fp32 CUDA tensors (srd_axpby)."""
    xc, yc = x.contiguous().float(), y.contiguous().float()
    out = torch.empty_like(xc)
    with torch.cuda.device(xc.device):
        L.check(L.lib().srd_axpby(xc.data_ptr(), yc.data_ptr(), float(a), float(b), xc.numel(), out.data_ptr(),
                                  E._stream()), "srd_axpby")
    return out


def _run_single(x, build, ctx=None):
    """Stand-alone forward of a sub-module: NCHW fp32 in -> program -> NCHW fp32 out."""
    E._require_cuda(x, "input")
    x = x.contiguous().float()
    B, C_, H, W = x.shape
    with torch.cuda.device(x.device):          # plans, streams and launches belong to the tensor's device
        bld = E.Builder(x.device, runtime.precision)
        out_act, out_c = build(bld, x, B, C_, H, W)
        out = torch.empty((B, out_c, out_act.H, out_act.W), dtype=torch.float32, device=x.device)
        bld.act_to_nchw(out_act, 0, out_c, out)
        bld.run(ctx)
    return out


class Block(nn.Module):
    """:439-457.  groups == 0 (what the Unet hard-wires, :757): ReflectionPad2d, Conv2d, Mish -- one fused launch.
    groups > 0: a GroupNorm between conv and Mish (:448-454) -- bias-only conv launch + srd_groupnorm_act."""

    def __init__(self, dim, dim_out, groups=8):
        super().__init__()
        if groups == 0:
            self.block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim_out, 3), Mish())
        else:
            self.block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim_out, 3),
                                       nn.GroupNorm(groups, dim_out), Mish())

    def forward(self, x):
        conv = self.block[1]

        def build(bld, x, B, C_, H, W):
            xin = bld.act(B, H, W, C_)
            bld.nchw_to_act(lambda: x, xin, C_, border=L.BORDER_REFLECT)
            out = bld.act(B, H, W, conv.out_channels)
            if len(self.block) == 4:
                tmp = bld.act(B, H, W, conv.out_channels)
                bld.conv(E._pc(conv), [(xin, 0, C_)], out=tmp, act=L.ACT_NONE)
                bld.groupnorm(tmp, conv.out_channels, self.block[2], out, act=L.ACT_MISH)
            else:
                bld.conv(E._pc(conv), [(xin, 0, C_)], out=out, act=L.ACT_MISH)
            return out, conv.out_channels
        return _run_single(x, build)


class ResnetBlock(nn.Module):
    """:460-480"""

    def __init__(self, dim, dim_out, *, time_emb_dim=0, groups=8):
        super().__init__()
        if time_emb_dim > 0:
            self.mlp = nn.Sequential(Mish(), nn.Linear(time_emb_dim, dim_out))
        self.block1 = Block(dim, dim_out, groups=groups)
        self.block2 = Block(dim_out, dim_out, groups=groups)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self._dim, self._dim_out = dim, dim_out

    def forward(self, x, time_emb=None, cond=None):
        """:473-480.  Stand-alone call: `time_emb` [B, time_emb_dim] goes through `self.mlp` (Mish + Linear, one
        srd_linear_f32 launch) and is added per sample after block1; `cond` [B, dim_out, H, W] likewise (:478).
        (Inside Unet.forward the same sum comes from the precomputed time table instead.)"""
        E._require_cuda(x, "input")
        temb = None
        if time_emb is not None and hasattr(self, "mlp"):
            te = time_emb.contiguous().float()
            lin = self.mlp[1]
            w, b = lin.weight.detach().float().contiguous(), lin.bias.detach().float().contiguous()
            temb = torch.empty((te.shape[0], lin.out_features), dtype=torch.float32, device=x.device)
            L.check(L.lib().srd_linear_f32(te.data_ptr(), w.data_ptr(), b.data_ptr(), te.shape[0], te.shape[1],
                                           lin.out_features, 1, temb.data_ptr(), E._stream()), "srd_linear_f32")
            if temb.shape[0] != x.shape[0]:
                raise RuntimeError("time_emb must have one row per image")

        def build(bld, x, B, C_, H, W):
            xin = bld.act(B, H, W, C_)
            bld.nchw_to_act(lambda: x, xin, C_, border=L.BORDER_REFLECT)
            mid = None
            if cond is not None:
                ca = bld.act(B, H, W, self._dim_out)
                bld.nchw_to_act(lambda: cond.contiguous().float(), ca, self._dim_out)
                mid = (ca, 0, 1.0)
            out = bld.act(B, H, W, self._dim_out)
            E.build_resnet(bld, self, [(xin, 0, C_)], self._dim_out, out, L.BORDER_KEEP,
                           0 if temb is not None else None, temb_bstride=self._dim_out, mid_res=mid)
            return out, self._dim_out
        return _run_single(x, build, E.RunCtx(temb.data_ptr()) if temb is not None else None)


class Residual(nn.Module):
    """:398-404.  Inside Unet.forward `fn(x) + x` is a residual epilogue of the launch list; the stand-alone
    forward runs `fn` (one of this module's CUDA-backed blocks) and adds with srd_axpby."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, *args, **kwargs):
        E._require_cuda(x, "Residual input")
        return _axpby(self.fn(x, *args, **kwargs), x, 1.0, 1.0)


class Rezero(nn.Module):
    """:427-434 -- `fn(x) * g`, g initialised to 0."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn
        self.g = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        E._require_cuda(x, "Rezero input")
        y = self.fn(x)
        return _axpby(y, y, float(self.g.detach().float().item()), 0.0)


class LinearAttention(nn.Module):
    """:506-522.  (Upstream's forward needs einops' `rearrange`, which the reference file never imports -- SURVEY Q6;
    the semantics built here are those of the code with that import present.)"""

    def __init__(self, dim, heads=4, dim_head=32):
        super().__init__()
        self.heads = heads
        hidden_dim = dim_head * heads
        self.to_qkv = nn.Conv2d(dim, hidden_dim * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden_dim, dim, 1)

    def forward(self, x):
        def build(bld, x, B, C_, H, W):
            hidden = self.to_out.in_channels
            xin = bld.act(B, H, W, C_)
            bld.nchw_to_act(lambda: x, xin, C_)
            qkv, ao, out = bld.act(B, H, W, 3 * hidden), bld.act(B, H, W, hidden), bld.act(B, H, W, C_)
            bld.conv(E._pc(self.to_qkv), [(xin, 0, C_)], out=qkv)
            bld.linear_attention(qkv, self.heads, hidden // self.heads, ao)
            bld.conv(E._pc(self.to_out), [(ao, 0, hidden)], out=out)
            return out, C_
        return _run_single(x, build)


class Upsample(nn.Module):
    """:483-491 ConvTranspose2d(dim, dim, 4, 2, 1)"""

    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Sequential(nn.ConvTranspose2d(dim, dim, 4, 2, 1))

    def forward(self, x):
        conv = self.conv[0]

        def build(bld, x, B, C_, H, W):
            xin = bld.act(B, H, W, C_)
            bld.nchw_to_act(lambda: x, xin, C_, border=L.BORDER_KEEP)
            out = bld.act(B, 2 * H, 2 * W, conv.out_channels)
            bld.conv(E._pc(conv), [(xin, 0, C_)], out=out)
            return out, conv.out_channels
        return _run_single(x, build)


class Downsample(nn.Module):
    """:494-503 ReflectionPad2d(1) + Conv2d(dim, dim, 3, 2)"""

    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3, 2))

    def forward(self, x):
        conv = self.conv[1]

        def build(bld, x, B, C_, H, W):
            xin = bld.act(B, H, W, C_)
            bld.nchw_to_act(lambda: x, xin, C_, border=L.BORDER_REFLECT)
            out = bld.act(B, H // 2, W // 2, conv.out_channels)
            bld.conv(E._pc(conv), [(xin, 0, C_)], out=out)
            return out, conv.out_channels
        return _run_single(x, build)


# ------------------------------------------------------------------------------------------ RRDB
def make_layer(block, n_layers, seq=False):
    """:98-105"""
    layers = [block() for _ in range(n_layers)]
    return nn.Sequential(*layers) if seq else nn.ModuleList(layers)


class ResidualDenseBlock_5C(nn.Module):
    """:670-690 (parameter container; executed by engine.RRDBProgram)."""

    def __init__(self, nf=64, gc=32, bias=True):
        super().__init__()
        self.conv1 = nn.Conv2d(nf, gc, 3, 1, 1, bias=bias)
        self.conv2 = nn.Conv2d(nf + gc, gc, 3, 1, 1, bias=bias)
        self.conv3 = nn.Conv2d(nf + 2 * gc, gc, 3, 1, 1, bias=bias)
        self.conv4 = nn.Conv2d(nf + 3 * gc, gc, 3, 1, 1, bias=bias)
        self.conv5 = nn.Conv2d(nf + 4 * gc, nf, 3, 1, 1, bias=bias)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        self._nf, self._gc = nf, gc

    def forward(self, x):
        """:680-690 stand-alone: the four growth convs write channel slices of ONE (nf + 4 gc)-channel buffer (the
        `cat`s), conv5's epilogue carries `x5 * 0.2 + x`.  (Inside RRDBNet the same launches come from
        engine.RRDBProgram.)"""
        def build(bld, x, B, C_, H, W):
            out = bld.act(B, H, W, self._nf)
            cat = E.build_rdb_input(bld, x, self._nf, self._gc)
            E.build_rdb(bld, self, cat, self._nf, self._gc, out, alpha=0.2, res1=(cat, 0, 1.0))
            return out, self._nf
        return _run_single(x, build)


class RRDB(nn.Module):
    """:693-706"""

    def __init__(self, nf, gc=32):
        super().__init__()
        self.RDB1 = ResidualDenseBlock_5C(nf, gc)
        self.RDB2 = ResidualDenseBlock_5C(nf, gc)
        self.RDB3 = ResidualDenseBlock_5C(nf, gc)
        self._nf, self._gc = nf, gc

    def forward(self, x):
        """:702-706 stand-alone: three dense blocks, then `out * 0.2 + x` folded into the last conv5's epilogue
        (alpha 0.04, 0.2 * RDB3-input, 1.0 * x) exactly as engine.RRDBProgram does."""
        def build(bld, x, B, C_, H, W):
            nf, gc = self._nf, self._gc
            cats = [E.build_rdb_input(bld, x, nf, gc), bld.act(B, H, W, nf + 4 * gc)]
            keep = bld.act(B, H, W, nf)
            bld.nchw_to_act(lambda: x, keep, nf)
            out = bld.act(B, H, W, nf)
            E.build_rdb(bld, self.RDB1, cats[0], nf, gc, cats[1], alpha=0.2, res1=(cats[0], 0, 1.0))
            E.build_rdb(bld, self.RDB2, cats[1], nf, gc, cats[0], alpha=0.2, res1=(cats[1], 0, 1.0))
            E.build_rdb(bld, self.RDB3, cats[0], nf, gc, out, alpha=0.04, res1=(cats[0], 0, 0.2), res2=(keep, 0, 1.0))
            return out, nf
        return _run_single(x, build)


class RRDBNet(_EngineCached):
    _cache_fields = ("_progs",)

    """:709-749"""

    def __init__(self, in_nc, out_nc, nf, nb, gc=32):
        super().__init__()
        self.conv_first = nn.Conv2d(in_nc, nf, 3, 1, 1, bias=True)
        self.RRDB_trunk = make_layer(partial(RRDB, nf=nf, gc=gc), nb)
        self.trunk_conv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv1 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.upconv2 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        if hparams['sr_scale'] == 8:
            self.upconv3 = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.HRconv = nn.Conv2d(nf, nf, 3, 1, 1, bias=True)
        self.conv_last = nn.Conv2d(nf, out_nc, 3, 1, 1, bias=True)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)
        self._in_nc, self._out_nc, self._nf, self._gc = in_nc, out_nc, nf, gc
        self._progs = {}

    def _program(self, B, h, w, device):
        key = (B, h, w, runtime.precision, str(device), E.params_signature(self), hparams['sr_scale'])
        p = self._progs.get("p")
        if p is None or p[0] != key:
            p = (key, E.RRDBProgram(self, B, h, w, runtime.precision, device))
            self._progs["p"] = p
        return p[1]

    @torch.no_grad()
    def forward(self, x, get_fea=False):
        E._require_cuda(x, "RRDBNet input")
        x = x.contiguous().float()
        B, _, h, w = x.shape
        with torch.cuda.device(x.device):
            prog = self._program(B, h, w, x.device)
            prog.load_input(x)
            prog.run()
            out = prog.out.clone()
            if get_fea:
                return out, prog.feas_nchw()
            return out


# ------------------------------------------------------------------------------------------ Unet
class Unet(_EngineCached):
    _cache_fields = ("_progs",)

    """:752-856"""

    def __init__(self, dim, out_dim=None, dim_mults=(1, 2, 4, 8), cond_dim=32):
        super().__init__()
        dims = [3, *map(lambda m: dim * m, dim_mults)]
        in_out = list(zip(dims[:-1], dims[1:]))
        groups = 0
        self.cond_proj = nn.ConvTranspose2d(cond_dim * ((hparams['rrdb_num_block'] + 1) // 3), dim,
                                            hparams['sr_scale'] * 2, hparams['sr_scale'], hparams['sr_scale'] // 2)
        self.time_pos_emb = SinusoidalPosEmb(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, dim * 4), Mish(), nn.Linear(dim * 4, dim))
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        num_resolutions = len(in_out)
        for ind, (dim_in, dim_out) in enumerate(in_out):
            is_last = ind >= (num_resolutions - 1)
            self.downs.append(nn.ModuleList([
                ResnetBlock(dim_in, dim_out, time_emb_dim=dim, groups=groups),
                ResnetBlock(dim_out, dim_out, time_emb_dim=dim, groups=groups),
                Downsample(dim_out) if not is_last else nn.Identity()]))
        mid_dim = dims[-1]
        self.mid_block1 = ResnetBlock(mid_dim, mid_dim, time_emb_dim=dim, groups=groups)
        if hparams['use_attn']:
            self.mid_attn = Residual(Rezero(LinearAttention(mid_dim)))
        self.mid_block2 = ResnetBlock(mid_dim, mid_dim, time_emb_dim=dim, groups=groups)
        for ind, (dim_in, dim_out) in enumerate(reversed(in_out[1:])):
            is_last = ind >= (num_resolutions - 1)          # never true: three Upsamples
            self.ups.append(nn.ModuleList([
                ResnetBlock(dim_out * 2, dim_in, time_emb_dim=dim, groups=groups),
                ResnetBlock(dim_in, dim_in, time_emb_dim=dim, groups=groups),
                Upsample(dim_in) if not is_last else nn.Identity()]))
        self.final_conv = nn.Sequential(Block(dim, dim, groups=groups), nn.Conv2d(dim, out_dim, 1))
        if hparams['res'] and hparams['up_input']:
            self.up_proj = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(3, dim, 3))      # :803-806
        if hparams['use_wn']:
            self.apply_weight_norm()
        if hparams['weight_init']:
            self.apply(initialize_weights)
        self._dim, self._dims, self._out_dim, self._cond_dim = dim, dims, out_dim, cond_dim
        self._progs = {}

    def apply_weight_norm(self):
        """:808-814: weight_norm on every Conv1d / Conv2d (not the transposed convs, not the Linears) -- the state
        dict then carries `weight_g` / `weight_v`; the engine folds them when it packs (engine.eff_weight)."""
        def _apply_weight_norm(m):
            if isinstance(m, (torch.nn.Conv1d, torch.nn.Conv2d)):
                torch.nn.utils.weight_norm(m)
        self.apply(_apply_weight_norm)

    def make_generation_fast_(self):
        """:849-856"""
        def remove_weight_norm(m):
            try:
                nn.utils.remove_weight_norm(m)
            except ValueError:      # this module didn't have weight norm
                return
        self.apply(remove_weight_norm)

    def resnet_blocks(self):
        """Fixed order used for the columns of the time table."""
        blocks = []
        for r0, r1, _ in self.downs:
            blocks += [r0, r1]
        blocks += [self.mid_block1, self.mid_block2]
        for r0, r1, _ in self.ups:
            blocks += [r0, r1]
        return blocks

    @torch.no_grad()
    def forward(self, x, time, cond, img_lr_up):
        """x [B,3,H,W], time [B] (any per-sample values), cond = list of RRDB feature maps."""
        E._require_cuda(x, "Unet input")
        with torch.cuda.device(x.device):
            return self._forward(x, time, cond, img_lr_up)

    def _forward(self, x, time, cond, img_lr_up):
        x = x.contiguous().float()
        B, _, H, W = x.shape
        dev = x.device
        sel = cond[2::3]
        key = (B, H, W, runtime.precision, str(dev), E.params_signature(self),
               tuple(tuple(f.shape) for f in sel), bool(hparams['res'] and hparams['up_input']))
        p = self._progs.get("p")
        if p is None or p[0] != key:
            bld = E.Builder(dev, runtime.precision)
            fea_acts = [bld.act(B, f.shape[2], f.shape[3], f.shape[1]) for f in cond]
            if any(f.shape[1] != self._cond_dim for f in sel):
                raise RuntimeError("cond features have %d channels, Unet was built with cond_dim=%d"
                                   % (sel[0].shape[1], self._cond_dim))
            cprog = E.CondProgram(self, fea_acts, bld.prec, dev)
            uprog = E.UnetProgram(self, B, H, W, bld.prec, dev, cond_act=cprog.cond, per_sample_temb=True,
                                  dcond_act=cprog.dcond)
            p = (key, (bld, fea_acts, cprog, uprog))
            self._progs["p"] = p
        bld, fea_acts, cprog, uprog = p[1]
        for f, a in list(zip(cond, fea_acts))[2::3]:
            fc = f.contiguous().float()
            L.check(L.lib().srd_nchw_to_act(fc.data_ptr(), B, fc.shape[1], a.H, a.W, 1.0, 0.0, a.ptr, a.ct,
                                            bld.prec.dtype, L.BORDER_KEEP, E._stream()), "srd_nchw_to_act")
        if cprog.lr_up is not None:
            cprog.load_lr_up(img_lr_up.contiguous().float())
        cprog.run()
        L.check(L.lib().srd_nchw_to_act(x.data_ptr(), B, x.shape[1], H, W, 1.0, 0.0, uprog.xin.ptr, uprog.xin.ct,
                                        bld.prec.dtype, L.BORDER_REFLECT, E._stream()), "srd_nchw_to_act")
        table = E.time_table(self, time.reshape(-1).float())
        if table.shape[0] != B:
            raise RuntimeError("time must have one entry per image")
        self._last_table = table                         # must outlive the asynchronous launches
        uprog.run(table.data_ptr())
        return uprog.eps.clone()


# ------------------------------------------------------------------------------------------ schedule
def extract(a, t, x_shape):
    """:860-863 (kept for API parity; on the hot path the gathers live inside srd_p_sample_step)."""
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    """:866-869"""
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


def _warmup_beta(beta_start, beta_end, num_diffusion_timesteps, warmup_frac):
    betas = beta_end * np.ones(num_diffusion_timesteps, dtype=np.float64)
    warmup_time = int(num_diffusion_timesteps * warmup_frac)
    betas[:warmup_time] = np.linspace(beta_start, beta_end, warmup_time, dtype=np.float64)
    return betas


def get_beta_schedule(num_diffusion_timesteps, beta_schedule='linear', beta_start=0.0001, beta_end=0.02):
    """:879-895"""
    T = num_diffusion_timesteps
    if beta_schedule == 'quad':
        betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=np.float64) ** 2
    elif beta_schedule == 'linear':
        betas = np.linspace(beta_start, beta_end, T, dtype=np.float64)
    elif beta_schedule == 'warmup10':
        betas = _warmup_beta(beta_start, beta_end, T, 0.1)
    elif beta_schedule == 'warmup50':
        betas = _warmup_beta(beta_start, beta_end, T, 0.5)
    elif beta_schedule == 'const':
        betas = beta_end * np.ones(T, dtype=np.float64)
    elif beta_schedule == 'jsd':
        betas = 1. / np.linspace(T, 1, T, dtype=np.float64)
    else:
        raise NotImplementedError(beta_schedule)
    assert betas.shape == (T,)
    return betas


def cosine_beta_schedule(timesteps, s=0.008):
    """:898-908 -- note linspace(0, steps, steps)."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas_cumprod = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


# ------------------------------------------------------------------------------------------ sampler
class GaussianDiffusion(_EngineCached):
    _cache_fields = ("_plans",)

    """:911-1134 (inference path)."""

    def __init__(self, denoise_fn, rrdb_net, timesteps=1000, loss_type='l1'):
        super().__init__()
        self.denoise_fn = denoise_fn
        self.rrdb = rrdb_net
        if hparams['beta_schedule'] == 'cosine':
            betas = cosine_beta_schedule(timesteps, s=hparams['beta_s'])
        if hparams['beta_schedule'] == 'linear':
            betas = get_beta_schedule(timesteps, beta_end=hparams['beta_end'])
            if hparams['res']:
                betas[-1] = 0.999
        alphas = 1. - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        to_torch = partial(torch.tensor, dtype=torch.float32)
        self.register_buffer('betas', to_torch(betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(alphas_cumprod_prev))
        self.register_buffer('sqrt_alphas_cumprod', to_torch(np.sqrt(alphas_cumprod)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - alphas_cumprod)))
        self.register_buffer('log_one_minus_alphas_cumprod', to_torch(np.log(1. - alphas_cumprod)))
        self.register_buffer('sqrt_recip_alphas_cumprod', to_torch(np.sqrt(1. / alphas_cumprod)))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', to_torch(np.sqrt(1. / alphas_cumprod - 1)))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        self.register_buffer('posterior_variance', to_torch(posterior_variance))
        self.register_buffer('posterior_log_variance_clipped',
                             to_torch(np.log(np.maximum(posterior_variance, 1e-20))))
        self.register_buffer('posterior_mean_coef1',
                             to_torch(betas * np.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod)))
        self.register_buffer('posterior_mean_coef2',
                             to_torch((1. - alphas_cumprod_prev) * np.sqrt(alphas) / (1. - alphas_cumprod)))
        self.sample_tqdm = True
        self._plans = {}
        self.last_stats = {}

    # ---- plan cache -------------------------------------------------------------------------
    def _plan(self, B, h, w, device):
        sig = E.plan_signature(self)
        key = (B, h, w, runtime.precision, str(device))
        p = self._plans.pop(key, None)
        if p is None or p.sig != sig:
            while len(self._plans) >= max(1, runtime.plan_cache_size):      # LRU: dicts keep insertion order
                self._plans.pop(next(iter(self._plans)))
            with torch.cuda.device(device):
                p = E.SamplerPlan(self, B, h, w, runtime.precision, device)
        self._plans[key] = p                                                # most recently used last
        return p

    def _elementwise(self, x, img_lr_up, clip_input, inverse):
        E._require_cuda(x, "image")
        if clip_input is None:
            clip_input = hparams['clip_input']
        if not hparams['res']:
            return x
        xc, uc = x.contiguous().float(), img_lr_up.contiguous().float()
        out = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            L.check(L.lib().srd_res2img(xc.data_ptr(), uc.data_ptr(), out.data_ptr(), xc.numel(),
                                        float(hparams['res_rescale']), 1 if clip_input else 0, inverse, E._stream()),
                    "srd_res2img")
        return out

    def res2img(self, img_, img_lr_up, clip_input=None):
        """:1118-1125"""
        return self._elementwise(img_, img_lr_up, clip_input, 0)

    def img2res(self, x, img_lr_up, clip_input=None):
        """:1127-1134"""
        return self._elementwise(x, img_lr_up, clip_input, 1)

    def _check_t(self, t, what, allow_negative=False):
        """The reference's `extract` gathers raise on an out-of-range timestep; the kernels index the coefficient
        tables directly, so the public entry points validate on the host (one tiny device->host read)."""
        tt = t.reshape(-1)
        if tt.numel() == 0:
            return
        lo, hi = int(tt.min()), int(tt.max())
        if hi >= self.num_timesteps or (lo < 0 and not allow_negative):
            raise IndexError("%s: timestep out of range [0, %d): min %d, max %d" % (what, self.num_timesteps, lo, hi))

    def _scale_pair(self, x, y, tab_a, tab_b, t, what):
        """tab_a[t] * x + tab_b[t] * y per sample (fp32 mul, mul, add -- the reference's op order) via srd_q_sample."""
        E._require_cuda(x, what)
        self._check_t(t, what)
        xc, yc = x.contiguous().float(), y.contiguous().float()
        t32 = t.to(device=xc.device, dtype=torch.int32).contiguous()
        out = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            L.check(L.lib().srd_q_sample(xc.data_ptr(), yc.data_ptr(), tab_a.data_ptr(), tab_b.data_ptr(),
                                         t32.data_ptr(), xc.shape[0], xc[0].numel(), out.data_ptr(), E._stream()),
                    "srd_q_sample")
        return out

    def _neg(self, name):
        """-buffer, cached (a*x - b*n == a*x + (-b)*n bit for bit: negation is exact)."""
        buf = getattr(self, name)
        key = (buf.data_ptr(), buf._version)
        hit = self.__dict__.setdefault("_srd_neg", {}).get(name)
        if hit is None or hit[0] != key:
            hit = (key, (-buf.detach()).contiguous())
            self.__dict__["_srd_neg"][name] = hit
        return hit[1]

    def q_mean_variance(self, x_start, t):
        """:962-966"""
        zeros = torch.zeros_like(self.sqrt_alphas_cumprod)
        mean = self._scale_pair(x_start, x_start, self.sqrt_alphas_cumprod, zeros, t, "q_mean_variance input")
        variance = extract(1. - self.alphas_cumprod, t, x_start.shape)
        log_variance = extract(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def predict_start_from_noise(self, x_t, t, noise):
        """:968-972"""
        return self._scale_pair(x_t, noise, self.sqrt_recip_alphas_cumprod, self._neg("sqrt_recipm1_alphas_cumprod"),
                                t, "predict_start_from_noise input")

    def q_posterior(self, x_start, x_t, t):
        """:974-981"""
        posterior_mean = self._scale_pair(x_start, x_t, self.posterior_mean_coef1, self.posterior_mean_coef2, t,
                                          "q_posterior input")
        posterior_variance = extract(self.posterior_variance, t, x_t.shape)
        posterior_log_variance_clipped = extract(self.posterior_log_variance_clipped, t, x_t.shape)
        return posterior_mean, posterior_variance, posterior_log_variance_clipped

    def p_mean_variance(self, x, t, noise_pred, clip_denoised: bool):
        """:983-990 -- the posterior mean and the (clamped) x_0 prediction come out of the fused p_sample kernel run
        with an all-zero noise tensor (mean + sigma * 0 == mean)."""
        E._require_cuda(x, "p_mean_variance input")
        self._check_t(t, "p_mean_variance")
        xc = x.contiguous().float()
        t32 = t.to(device=xc.device, dtype=torch.int32).contiguous()
        mean, x0 = torch.empty_like(xc), torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            E.p_sample_step(xc, noise_pred.contiguous().float(), torch.zeros_like(xc), E.coef_table(self), t32, mean,
                            clip_denoised, x0_out=x0)
        return (mean, extract(self.posterior_variance, t, xc.shape),
                extract(self.posterior_log_variance_clipped, t, xc.shape), x0)

    def forward(self, img_hr, img_lr, img_lr_up, t=None, *args, **kwargs):
        """:992-1019 is the TRAINING objective (q_sample -> denoiser -> p_losses with autograd); this package is the
        inference hot path only (training is out of scope), so calling the module raises instead of
        falling into nn.Module's default."""
        raise NotImplementedError(
            "dgm_img_super_resolution_b200.GaussianDiffusion is inference-only: use sample() / p_sample() / "
            "interpolate(); the training forward / p_losses of srdiff/model.py:992-1039 are out of scope")

    def p_losses(self, x_start, t, cond, img_lr_up, noise=None):
        """:1021-1039 -- training loss, out of scope (see forward)."""
        raise NotImplementedError("p_losses (training objective) is out of scope of the inference-only build")

    def q_sample(self, x_start, t, noise=None):
        """:1041-1048 forward diffusion of x_start to step t (per-sample t; t < 0 returns x_start)."""
        E._require_cuda(x_start, "q_sample input")
        self._check_t(t, "q_sample", allow_negative=True)
        x = x_start.contiguous().float()
        noise = default(noise, lambda: torch.randn_like(x)).contiguous().float()
        t32 = t.to(device=x.device, dtype=torch.int32).contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(L.lib().srd_q_sample(x.data_ptr(), noise.data_ptr(), self.sqrt_alphas_cumprod.data_ptr(),
                                         self.sqrt_one_minus_alphas_cumprod.data_ptr(), t32.data_ptr(), x.shape[0],
                                         x[0].numel(), out.data_ptr(), E._stream()), "srd_q_sample")
        return out

    @torch.no_grad()
    def interpolate(self, x1, x2, img_lr, img_lr_up, t=None, lam=0.5):
        """:1093-1116: diffuse x1 and x2 to step t, blend, and denoise from step t-1 down to 0."""
        E._require_cuda(x1, "interpolate input")
        with torch.cuda.device(x1.device):
            return self._interpolate(x1, x2, img_lr, img_lr_up, t, lam)

    def _interpolate(self, x1, x2, img_lr, img_lr_up, t, lam):
        if not hparams['use_rrdb']:
            raise NotImplementedError("hparams['use_rrdb']=False is outside the hot path")
        assert x1.shape == x2.shape
        b = x1.shape[0]
        device = x1.device
        t = default(t, self.num_timesteps - 1)
        B, _, h, w = img_lr.shape
        plan = self._plan(B, h, w, device)
        r1, r2 = self.img2res(x1, img_lr_up), self.img2res(x2, img_lr_up)
        t_batched = torch.full((b,), int(t), device=device, dtype=torch.long)
        xt1, xt2 = self.q_sample(r1, t=t_batched), self.q_sample(r2, t=t_batched)
        L.check(L.lib().srd_axpby(xt1.data_ptr(), xt2.data_ptr(), float(1 - lam), float(lam), xt1.numel(),
                                  plan.x.data_ptr(), E._stream()), "srd_axpby")
        plan.img_lr.copy_(img_lr)
        plan.img_lr_up.copy_(img_lr_up)
        seed = next_seed(device)
        plan.seed_dev.fill_(seed)
        plan.encode()
        plan.load_x()
        use_torch_noise = runtime.noise == "torch"
        for k, i in enumerate(reversed(range(0, int(t)))):
            noise = noise_like(plan.x.shape, device).contiguous().float() if use_torch_noise else None
            plan.step(i, k, noise, seed, runtime.image_offset, fuse_img=False)
        return self.res2img(plan.x.clone(), plan.img_lr_up)

    @torch.no_grad()
    def p_sample(self, x, t, cond, img_lr_up, noise_pred=None, clip_denoised=True, repeat_noise=False):
        """:1050-1060 stand-alone (arbitrary per-sample t).  Returns (x_prev, x0_pred)."""
        E._require_cuda(x, "p_sample input")
        self._check_t(t, "p_sample")
        x = x.contiguous().float()
        if noise_pred is None:
            noise_pred = self.denoise_fn(x, t, cond=cond, img_lr_up=img_lr_up)
        noise = noise_like(x.shape, x.device, repeat_noise).contiguous().float()
        coef = E.coef_table(self)
        t32 = t.to(device=x.device, dtype=torch.int32).contiguous()
        x_out, x0 = torch.empty_like(x), torch.empty_like(x)
        with torch.cuda.device(x.device):
            E.p_sample_step(x, noise_pred.contiguous().float(), noise, coef, t32, x_out, clip_denoised, x0_out=x0)
        return x_out, x0

    @torch.no_grad()
    def sample(self, img_lr, img_lr_up, shape, save_intermediate=False):
        """:1062-1091.  Returns (img, rrdb_out[, images])."""
        E._require_cuda(self.betas, "GaussianDiffusion")
        E._require_cuda(img_lr, "img_lr")
        with torch.cuda.device(self.betas.device):      # the module's device, not whatever happens to be current
            return self._sample(img_lr, img_lr_up, shape, save_intermediate)

    def _sample(self, img_lr, img_lr_up, shape, save_intermediate):
        device = self.betas.device
        if not hparams['use_rrdb']:
            raise NotImplementedError("hparams['use_rrdb']=False is outside the hot path")
        b = shape[0]
        B, _, h, w = img_lr.shape
        if b != B or tuple(shape) != tuple(img_lr_up.shape):
            raise RuntimeError("shape %s inconsistent with img_lr %s / img_lr_up %s"
                               % (tuple(shape), tuple(img_lr.shape), tuple(img_lr_up.shape)))
        plan = self._plan(B, h, w, device)
        if (plan.H, plan.W) != (shape[2], shape[3]):
            raise RuntimeError("img_lr_up must be sr_scale x img_lr")
        T = self.num_timesteps
        use_torch_noise = runtime.noise == "torch" and plan.noise_bytes() <= runtime.noise_budget_bytes
        seed = next_seed(device)
        # ---- x_T :1066-1070 (drawn before the RRDB forward, like the reference)
        res = bool(hparams['res'])
        if not res:         # start from the diffused bicubic image :1066-1068; res2img is the identity then
            t_last = torch.full((B,), T - 1, device=device, dtype=torch.long)
            plan.x.copy_(self.q_sample(img_lr_up, t_last))
        elif use_torch_noise:
            plan.x.copy_(torch.randn(shape, device=device))
        else:
            L.check(L.lib().srd_randn(plan.x.data_ptr(), B, plan.x[0].numel(), seed, 0, runtime.image_offset,
                                      E._stream()), "srd_randn")
        plan.img_lr.copy_(img_lr)
        plan.img_lr_up.copy_(img_lr_up)
        eager = save_intermediate or not runtime.graph or not res
        images = []
        if use_torch_noise and not eager:
            nb = plan.ensure_noise_buf()
            for k in range(T):                       # one draw per step, t == 0 included
                nb[k].copy_(noise_like(shape, device))
        if not eager:
            plan.run_graph(use_torch_noise, seed, runtime.image_offset)
        else:
            plan.seed_dev.fill_(seed)
            plan.encode()
            plan.load_x()
            it = reversed(range(T))
            if self.sample_tqdm:
                from tqdm import tqdm
                it = tqdm(it, desc='sampling loop time step', total=T)
            x0 = torch.empty_like(plan.x) if save_intermediate else None
            for k, i in enumerate(it):
                noise = noise_like(shape, device).contiguous().float() if use_torch_noise else None
                plan.step(i, k, noise, seed, runtime.image_offset, x0_out=x0, fuse_img=res)
                if save_intermediate:
                    images.append((self.res2img(plan.x, plan.img_lr_up).cpu(),
                                   self.res2img(x0, plan.img_lr_up).cpu()))
        self.last_stats = {"gpu_launches": plan.n_launches(), "graph": not eager,
                           "noise": "torch" if use_torch_noise else "philox"}
        img = plan.img.clone() if res else plan.x.clone()
        rrdb_out = plan.rrdb_prog.out.clone()
        if save_intermediate:
            return img, rrdb_out, images
        return img, rrdb_out


# ------------------------------------------------------------------------------------------ checkpoints
# reading the upstream LeiaLi/SRDiff checkpoint format (:1405-1422, 1463-1496):
# `{work_dir}/model_ckpt_steps_{N}.ckpt` holding {'global_step', 'optimizer_states', 'state_dict': {'model': <state
# dict>}} (or a flat state dict with 'model.' prefixes).  Host-side file handling only; the state dict lands in the
# same nn.Module parameters the engine packs from.  The trainer's save / resume halves (:1425-1454: optimizer state,
# step counter) belong to the training loop, which is out of scope, and are not provided.
def get_all_ckpts(work_dir, steps=None):
    """:1416-1422 -- checkpoint paths, newest (largest step) first."""
    import glob
    import re
    pattern = f'{work_dir}/model_ckpt_steps_*.ckpt' if steps is None else f'{work_dir}/model_ckpt_steps_{steps}.ckpt'
    return sorted(glob.glob(pattern), key=lambda x: -int(re.findall(r'.*steps\_(\d+)\.ckpt', x)[0]))


def get_last_checkpoint(work_dir, steps=None):
    """:1405-1413"""
    checkpoint, last_ckpt_path = None, None
    ckpt_paths = get_all_ckpts(work_dir, steps)
    if len(ckpt_paths) > 0:
        last_ckpt_path = ckpt_paths[0]
        checkpoint = torch.load(last_ckpt_path, map_location='cpu')
    return checkpoint, last_ckpt_path


def load_ckpt(cur_model, ckpt_base_dir, model_name='model', force=True, strict=True):
    """:1463-1496 -- load `model_name`'s weights from a checkpoint file or the newest one in a directory."""
    import os
    if os.path.isfile(ckpt_base_dir):
        base_dir = os.path.dirname(ckpt_base_dir)
        ckpt_path = ckpt_base_dir
        checkpoint = torch.load(ckpt_base_dir, map_location='cpu')
    else:
        base_dir = ckpt_base_dir
        checkpoint, ckpt_path = get_last_checkpoint(ckpt_base_dir)
    if checkpoint is None:
        e_msg = f"| ckpt not found in {base_dir}."
        if force:
            assert False, e_msg
        print(e_msg)
        return
    state_dict = checkpoint["state_dict"]
    if len([k for k in state_dict.keys() if '.' in k]) > 0:
        state_dict = {k[len(model_name) + 1:]: v for k, v in state_dict.items() if k.startswith(f'{model_name}.')}
    else:
        state_dict = state_dict[model_name]
    if not strict:
        cur = cur_model.state_dict()
        unmatched = [k for k, p in state_dict.items() if k in cur and cur[k].shape != p.shape]
        for k in unmatched:
            print("| Unmatched keys: ", k, cur[k].shape, state_dict[k].shape)
            del state_dict[k]
    cur_model.load_state_dict(state_dict, strict=strict)
    print(f"| load '{model_name}' from '{ckpt_path}'.")
