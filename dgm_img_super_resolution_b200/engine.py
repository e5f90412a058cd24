"""Host-side engine: weight repacking, activation arena, layer programs and CUDA-graph capture.

PyTorch is used here only for device memory (torch.empty/zeros), streams, CUDA-graph capture
and one-time weight re-layout; every per-image / per-step FLOP runs in libsrdiff_b200.so
(see include/srdiff_b200.h).  Reference line numbers refer to /root/reference/srdiff/model.py.
"""
import ctypes as C
import math
import os

import torch

from . import _lib as L

PRECISIONS = ("bf16", "fp32", "bf16_direct")
# developer switch for A/B measurements: res_conv as a second output of block1's launch (SrdConvDesc.N2)
FUSE_RES_CONV = os.environ.get("SRD_NO_FUSED_RES") != "1"
# widest block whose res_conv rides on block1's launch: 128 keeps two TMEM accumulator stages (2 x 2 x 128 columns);
# 192 / 256 outputs fit too (2 x N <= 512 columns, SRD_FUSE_RES_MAX_N=256: 42 -> 39 launches) but leave ONE stage, i.e.
# no MMA / epilogue overlap inside the launch -- measured 2.664 vs 2.677 ms per pass, inside the noise, so off
FUSE_RES_MAX_N = int(os.environ.get("SRD_FUSE_RES_MAX_N", "128"))
# 256-channel stride-1 convs as two 128-channel groups sharing each weight block between two tile pairs
# (SrdConvDesc.n_split, SRD_NSPLIT=1): halves the weight bytes streamed from L2 per MMA, doubles the activation boxes;
# measured neutral on B200 (layers 16-24 of the Unet: 300-326 us vs 285-322 us, run-to-run spread larger than the
# difference), so off by default
NSPLIT = os.environ.get("SRD_NSPLIT", "0") == "1"
REVERSE_ORDER = os.environ.get("SRD_NO_REVERSE", "0") != "1"   # alternate the tile order of consecutive conv launches
FOLD_COND = os.environ.get("SRD_NO_FOLD_COND", "0") != "1"     # add cond after the first Downsample (see CondProgram)
# Unet.forward as ONE persistent launch (srd_chain_*), opt-in with SRD_CHAIN=1: bit-identical to the per-layer
# launches but measured SLOWER on B200 (its role loops lose the constant-bank / uniform-datapath operands
# of the single-layer kernel), so one launch per layer stays the default.  SRD_CHAIN_SINGLE=1 additionally drains the
# CTA pair after every layer (debug).
USE_CHAIN = os.environ.get("SRD_CHAIN", "0") == "1"
# Runs of consecutive layers with the same kernel plan (ResnetBlock pairs, the 20x20 level ...) as ONE fused-segment
# launch each (srd_chain_create flags bit 1: compile-time unrolled layer loop, constant-bank arguments kept), opt-in
# with SRD_FUSE=1: bit-identical, 42 -> 27 launches per pass, but each fused layer runs 5-20 % slower than its own
# launch (the dependency / publishing state costs the register-capped epilogue its software pipelining), which is more
# than the ~7 us per saved launch boundary -- measured 3.08 vs 2.83 ms per pass.
FUSE_SEGMENTS = os.environ.get("SRD_FUSE", "0") == "1"
CHAIN_SINGLE_SEGMENTS = os.environ.get("SRD_CHAIN_SINGLE", "0") == "1"


def _ceil(a, b):
    return (a + b - 1) // b


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise L.SrdError(
            "%s is on '%s': dgm_img_super_resolution_b200 runs only on CUDA (sm_100a); there is no CPU "
            "fallback. Move the module and its inputs to the GPU." % (what, t.device))


class Precision:
    def __init__(self, name):
        if name not in PRECISIONS:
            raise ValueError("precision must be one of %s" % (PRECISIONS,))
        self.name = name
        self.dtype = L.F32 if name == "fp32" else L.BF16
        self.torch_dtype = torch.float32 if name == "fp32" else torch.bfloat16
        self.kernel = L.KERNEL_AUTO if name == "bf16" else L.KERNEL_DIRECT


# --------------------------------------------------------------------------------------------
# activations: padded NHWC
# --------------------------------------------------------------------------------------------
class Act:
    """[B][H+2][W+2][ct] buffer; ring zero-initialised (consumers that need zero padding rely on it)."""

    def __init__(self, B, H, W, ct, prec, device, slack=0):
        self.B, self.H, self.W, self.ct = B, H, W, ct
        n = B * (H + 2) * (W + 2) * ct
        # `slack` extra zero elements after the last pixel (SRD_SRC_WIN8 windows read past the row end)
        self._flat = torch.zeros(n + slack, dtype=prec.torch_dtype, device=device)
        self.t = self._flat[:n].view(B, H + 2, W + 2, ct)

    @property
    def ptr(self):
        return self.t.data_ptr()


# --------------------------------------------------------------------------------------------
# weight packing (one-time re-layout; see include/srdiff_b200.h for the layouts)
# --------------------------------------------------------------------------------------------
def _host(weight):
    """fp32 host copy of a (small) parameter: the one-time re-layout runs as CPU tensor ops followed by ONE
    host->device copy per packed tensor, so building a plan launches no GPU kernels of its own (the first kernels
    of a fresh process are then the conv launches themselves, not hundreds of ATen slice copies)."""
    return weight.detach().to(device="cpu", dtype=torch.float32)


def _chunked(wp, nch):
    """[rows][nch*64][...] -> [rows][nch][64][...]"""
    return wp.reshape(wp.shape[0], nch, 64, *wp.shape[2:])


def _pack_conv_blocks(w, src_cins, ky_order):
    """Per source the K-blocks [nch][kw][ky*][Npad][64] (fp32) of an nn.Conv2d weight [O][I][kh][kw]: block
    (chunk, kx, ky) holds w[:, chunk*64 + c, ky, kx] as an [Npad][64] K-major tile, zero padded.  A handful of
    whole-tensor ops per source (pad, gather, permute) -- not one copy per (chunk, tap)."""
    O, I, kh, kw = w.shape
    npad = _ceil(O, 16) * 16
    ky_idx = torch.tensor(ky_order)
    out, a = [], 0
    for cin in src_cins:
        nch = _ceil(cin, 64)
        wp = torch.nn.functional.pad(w[:, a:a + cin], (0, 0, 0, 0, 0, nch * 64 - cin, 0, npad - O))
        blk = _chunked(wp, nch).index_select(3, ky_idx).permute(1, 4, 3, 0, 2)    # [nch][kw][ky*][npad][64]
        out.append(blk)
        a += cin
    return out


# tap order of a stride-2 conv's k-blocks when the kernel reads halo boxes per (row parity, column parity)
# (srd_tc_s2_halo(), csrc/conv_tc.cuh mma_role<TAPS_S2>): boxes (even, even), (even, odd), (odd, even), (odd, odd)
_S2_TAPS = [(0, 0), (0, 2), (2, 0), (2, 2), (0, 1), (2, 1), (1, 0), (1, 2), (1, 1)]


def _pack_conv_blocks_taps(w, src_cins, taps):
    """Like _pack_conv_blocks with an explicit (ky, kx) list: per source [nch][tap][Npad][64]."""
    O, I, kh, kw = w.shape
    npad = _ceil(O, 16) * 16
    idx = torch.tensor([ky * kw + kx for ky, kx in taps])
    out, a = [], 0
    for cin in src_cins:
        nch = _ceil(cin, 64)
        wp = torch.nn.functional.pad(w[:, a:a + cin], (0, 0, 0, 0, 0, nch * 64 - cin, 0, npad - O))
        blk = _chunked(wp, nch).reshape(npad, nch, 64, kh * kw).index_select(3, idx).permute(1, 3, 0, 2)
        out.append(blk)
        a += cin
    return out


def _pack_conv(weight, src_cins, stride):
    """nn.Conv2d weight [O][I][kh][kw] -> (w_direct fp32 [1][K][O], w_tc bf16 [KB][Npad][64])."""
    O, I, kh, kw = weight.shape
    assert sum(src_cins) == I, (src_cins, I)
    npad = _ceil(O, 16) * 16
    dev = weight.device
    w = _host(weight)
    ky_order = [0, 2, 1] if stride == 2 else list(range(kh))
    parts, a = [], 0
    for cin in src_cins:
        parts.append(w[:, a:a + cin].permute(2, 3, 1, 0).reshape(kh * kw * cin, O))
        a += cin
    w_direct = torch.cat(parts, 0)[None].contiguous()
    if stride == 2 and L.lib().srd_tc_s2_halo():
        blocks = [b.reshape(-1, npad, 64) for b in _pack_conv_blocks_taps(w, src_cins, _S2_TAPS)]
    else:
        blocks = [b.reshape(-1, npad, 64) for b in _pack_conv_blocks(w, src_cins, ky_order)]
    w_tc = torch.cat(blocks, 0).to(torch.bfloat16).contiguous()
    return w_direct.to(dev), w_tc.to(dev)


def _pack_convT(weight, src_cins, S):
    """nn.ConvTranspose2d(k=2S, stride=S, pad=S/2) weight [I][O][2S][2S] -> per sub-pixel phase
    (r, s) a 2x2-tap conv: out[y*S+r] += in[y+dy] * w[ky], ky = r + S/2 - dy*S (dy in {-1,0} for
    r < S/2, {0,1} otherwise).  SURVEY 7.4 tap tables.  All phases are gathered at once through the
    [S][2] tap-index table."""
    I, O, k, _ = weight.shape
    assert k == 2 * S and sum(src_cins) == I
    npad = _ceil(O, 16) * 16
    dev = weight.device
    w = _host(weight)
    taps = torch.tensor([[r + S // 2 - ((1 if r >= S // 2 else 0) + d - 1) * S for d in range(2)] for r in range(S)],
                        )                                                  # [S][2]: kernel index of tap d in phase r
    directs, tcs, a = [], [], 0
    for cin in src_cins:
        ws = w[a:a + cin]                                                  # [cin][O][k][k]
        g = ws[:, :, taps][..., taps]                                      # [cin][O][r][dy][s][dx]
        directs.append(g.permute(2, 4, 3, 5, 0, 1).reshape(S * S, 4 * cin, O))         # K order (dy, dx, c)
        nch = _ceil(cin, 64)
        gp = torch.nn.functional.pad(g, (0, 0, 0, 0, 0, 0, 0, 0, 0, npad - O, 0, nch * 64 - cin))
        gp = gp.reshape(nch, 64, npad, S, 2, S, 2)                         # [ch][c][n][r][dy][s][dx]
        tcs.append(gp.permute(3, 5, 0, 6, 4, 2, 1).reshape(S * S, nch * 4, npad, 64))  # blocks (chunk, dx, dy)
        a += cin
    w_direct = torch.cat(directs, 1).contiguous()
    w_tc = torch.cat(tcs, 1).to(torch.bfloat16).contiguous()
    return w_direct.to(dev), w_tc.to(dev)


def _pack_conv_dual(w3, w1, src_cins):
    """w_tc for a launch with a second 1x1 output (SrdConvDesc.N2): per (source, 64-channel chunk) the nine
    3x3 k-blocks in (kx, ky) order, then ONE k-block of the 1x1 weights [O][I][1][1] for the same channels."""
    O, I, kh, kw = w3.shape
    assert (kh, kw) == (3, 3) and tuple(w1.shape) == (O, I, 1, 1) and sum(src_cins) == I
    npad = _ceil(O, 16) * 16
    b3 = _pack_conv_blocks(_host(w3), src_cins, [0, 1, 2])
    b1 = _pack_conv_blocks(_host(w1), src_cins, [0])
    blocks = [torch.cat([x.reshape(x.shape[0], 9, npad, 64), y.reshape(y.shape[0], 1, npad, 64)], 1)
              for x, y in zip(b3, b1)]
    return torch.cat([b.reshape(-1, npad, 64) for b in blocks], 0).to(torch.bfloat16).contiguous().to(w3.device)


def _pack_conv_win8(weight):
    """3x3 nn.Conv2d weight [O][I<=8][3][3] for an SRD_SRC_WIN8 source: one K = 64 block per ky whose K
    index is dx*8 + c (dx = pixel inside the 8-pixel window = kx for dx < 3, zero weights beyond)."""
    O, I, kh, kw = weight.shape
    assert kh == 3 and kw == 3 and I <= 8
    npad = _ceil(O, 16) * 16
    w = _host(weight)
    blk = torch.zeros((kh, npad, 8, 8), dtype=torch.float32)                          # [ky][n][dx][c]
    blk[:, :O, :kw, :I] = w.permute(2, 0, 3, 1)                                       # [ky][O][kx][I]
    return blk.reshape(kh, npad, 64).to(torch.bfloat16).contiguous().to(weight.device)


class PackedConv:
    """Caches the packed forms of one conv layer's parameters, keyed by parameter identity/version."""

    def __init__(self, weight, bias, transposed=False, stride=1):
        self.weight, self.bias_p = weight, bias
        self.transposed, self.stride = transposed, stride
        self._cache = {}

    @property
    def N(self):
        return self.weight.shape[1] if self.transposed else self.weight.shape[0]

    @property
    def ksize(self):
        return self.weight.shape[2]

    def get_dual(self, src_cins, w1, b1):
        """(w_direct, w_tc with the interleaved 1x1 blocks, bias, bias2) -- see _pack_conv_dual."""
        key = ("dual", tuple(src_cins), self.weight.data_ptr(), self.weight._version, self.bias_p.data_ptr(),
               self.bias_p._version, w1.data_ptr(), w1._version, b1.data_ptr(), b1._version)
        hit = self._cache.get("du")
        if hit is None or hit[0] != key:
            wd, _ = _pack_conv(self.weight, src_cins, 1)
            self._cache["du"] = (key, wd, _pack_conv_dual(self.weight, w1, src_cins),
                                 self.bias_p.detach().float().contiguous(), b1.detach().float().contiguous())
            hit = self._cache["du"]
        return hit[1], hit[2], hit[3], hit[4]

    def get_win8(self):
        """(w_direct, w_tc_win8, bias) for an SRD_SRC_WIN8 source (see _pack_conv_win8)."""
        key = ("win8", self.weight.data_ptr(), self.weight._version, self.bias_p.data_ptr(), self.bias_p._version)
        hit = self._cache.get("w8")
        if hit is None or hit[0] != key:
            wd, _ = _pack_conv(self.weight, [self.weight.shape[1]], 1)
            self._cache["w8"] = (key, wd, _pack_conv_win8(self.weight), self.bias_p.detach().float().contiguous())
            hit = self._cache["w8"]
        return hit[1], hit[2], hit[3]

    def get(self, src_cins, n_split=1):
        key = (tuple(src_cins), n_split, self.weight.data_ptr(), self.weight._version,
               self.bias_p.data_ptr(), self.bias_p._version)
        slot = "k%d" % n_split
        hit = self._cache.get(slot)
        if hit is None or hit[0] != key:
            if self.transposed:
                wd, wt = _pack_convT(self.weight, src_cins, self.ksize // 2)
            else:
                wd, wt = _pack_conv(self.weight, src_cins, self.stride)
                if n_split > 1:       # w_tc per output-channel group: [group][KB][N / n_split][64]
                    O = self.weight.shape[0]
                    assert O % (64 * n_split) == 0
                    h = O // n_split
                    wt = torch.stack([_pack_conv(self.weight[g * h:(g + 1) * h], src_cins, self.stride)[1]
                                      for g in range(n_split)]).contiguous()
            b = self.bias_p.detach().float().contiguous()
            self._cache[slot] = (key, wd, wt, b)
            hit = self._cache[slot]
        return hit[1], hit[2], hit[3]


# --------------------------------------------------------------------------------------------
# ops
# --------------------------------------------------------------------------------------------
class ConvOp:
    def __init__(self, desc, keep, temb_off=None, io=None):
        self.desc, self.keep, self.temb_off = desc, keep, temb_off
        self.io = io or {}      # Act objects read / written (srcs, res1, res2, lin, out, out2, dst2): chain dependencies
        h = C.c_void_p()
        L.check(L.lib().srd_conv_plan_create(C.byref(desc), C.byref(h)), "srd_conv_plan_create")
        self.handle = h
        self.kernel = L.lib().srd_conv_plan_kernel(h)

    def run(self, ctx):
        temb = None
        if self.temb_off is not None:
            temb = C.c_void_p(ctx.temb_ptr + 4 * self.temb_off)
        L.check(L.lib().srd_conv_run(self.handle, temb, _stream()), "srd_conv_run")

    def rebuild(self):
        """Re-plan after a change of the descriptor (e.g. the tile order)."""
        h = C.c_void_p()
        L.check(L.lib().srd_conv_plan_create(C.byref(self.desc), C.byref(h)), "srd_conv_plan_create")
        L.lib().srd_conv_plan_destroy(self.handle)
        self.handle = h

    def __del__(self):
        try:
            if self.handle:
                L.lib().srd_conv_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class ChainOp:
    """A run of tcgen05 conv launches executed as ONE persistent kernel (include/srdiff_b200.h: srd_chain_*).
    Dependencies between the layers are derived from the activation buffers they read and write."""

    def __init__(self, ops, fused=False):
        self.ops = ops
        self.fused = fused
        n = len(ops)
        writer = {}                                    # id(Act) -> (layer, which output)
        links = (L.SrdChainLink * n)()
        for i, op in enumerate(ops):
            io, k = op.io, links[i]
            for s_ in range(L.SRD_MAX_SRC):
                k.src_layer[s_], k.src_which[s_] = -1, 0
            for s_, a in enumerate(io["srcs"]):
                k.src_layer[s_], k.src_which[s_] = writer.get(id(a), (-1, 0))
            k.res1_layer, k.res1_which = writer.get(id(io["res1"]), (-1, 0)) if io["res1"] is not None else (-1, 0)
            k.temb_off = op.temb_off if op.temb_off is not None else -1
            for key, which in (("out", 0), ("dst2", 1)):
                if io[key] is not None:
                    writer[id(io[key])] = (i, which)
        self.links = links
        handles = (C.c_void_p * n)(*[op.handle for op in ops])
        h = C.c_void_p()
        flags = 2 if fused else (1 if CHAIN_SINGLE_SEGMENTS else 0)
        L.check(L.lib().srd_chain_create(handles, links, n, flags, C.byref(h)), "srd_chain_create")
        self.handle = h
        self.n_segments = L.lib().srd_chain_segments(h)

    @staticmethod
    def chainable(ops):
        """Every op a tcgen05 conv with a batch-uniform time embedding; every buffer written once, never read before
        it is written, nothing in place; res2 / lin sources and second NHWC copies come from outside the chain."""
        written, read = set(), set()
        for op in ops:
            if not isinstance(op, ConvOp) or op.kernel != L.KERNEL_TC or op.desc.temb_bstride != 0:
                return False
            io = op.io
            if io.get("out2") is not None or op.desc.up not in (1, 2):
                return False
            ins = [id(a) for a in io["srcs"]] + [id(io[k]) for k in ("res1", "res2", "lin") if io[k] is not None]
            for k in ("res2", "lin"):
                if io[k] is not None and id(io[k]) in written:
                    return False
            outs = [id(io[k]) for k in ("out", "dst2") if io[k] is not None]
            if any(o in written or o in read or o in ins for o in outs):
                return False
            read.update(ins)
            written.update(outs)
        return 2 <= len(ops) <= L.lib().srd_chain_max_layers()

    def run(self, ctx):
        L.check(L.lib().srd_chain_run(self.handle, C.c_void_p(ctx.temb_ptr) if ctx.temb_ptr else None, _stream()),
                "srd_chain_run")

    def __del__(self):
        try:
            if self.handle:
                L.lib().srd_chain_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class FnOp:
    def __init__(self, fn, keep=()):
        self.fn, self.keep = fn, keep

    def run(self, ctx):
        self.fn(ctx)


class RunCtx:
    """Per-launch dynamic state: the address of the time-embedding row(s) for this step."""

    def __init__(self, temb_ptr=0):
        self.temb_ptr = temb_ptr


class Builder:
    """Accumulates an ordered list of kernel launches over a set of activation buffers."""

    def __init__(self, device, precision):
        self.device = device
        self.prec = precision if isinstance(precision, Precision) else Precision(precision)
        self.ops = []
        self.exec = None        # launch list after fuse_segments(): ConvOps and fused ChainOps
        self.chain = None       # ChainOp over all of `ops` (make_chain)
        self.chain_error = None
        self._dir = {}          # id(Act) -> 1 if its last writer walked the tiles in reverse order

    def act(self, B, H, W, C_, ct=None, slack=0):
        ct = _ceil(C_ if ct is None else ct, 8) * 8
        return Act(B, H, W, ct, self.prec, self.device, slack=slack)

    def run(self, ctx=None):
        ctx = ctx or RunCtx()
        if self.chain is not None:
            self.chain.run(ctx)
            return
        for op in (self.exec if self.exec is not None else self.ops):
            op.run(ctx)

    def fuse_segments(self):
        """Group runs of consecutive tcgen05 layers that share one kernel plan into fused-segment launches of at most
        srd_chain_fused_max_layers() layers (`self.exec` = what run() launches).  Inside a group every layer walks its
        tiles in the direction of the group's first layer: a consumer then needs the row blocks its producer finished
        FIRST, so it never waits for the producer's tail (across launches the serpentine order stays)."""
        self.exec = None
        if not FUSE_SEGMENTS or self.prec.name != "bf16":
            return
        lib, K = L.lib(), L.lib().srd_chain_fused_max_layers()
        ops, runs, cur = self.ops, [], []

        def plain(op):
            return (isinstance(op, ConvOp) and op.kernel == L.KERNEL_TC and op.desc.temb_bstride == 0 and
                    op.io.get("out2") is None and op.desc.up in (1, 2))
        for op in ops:
            if cur and plain(op) and lib.srd_conv_plans_share_segment(cur[-1].handle, op.handle) and \
                    ChainOp.chainable(cur + [op]):
                cur.append(op)
            else:
                if cur:
                    runs.append(cur)
                cur = [op] if plain(op) else []
                if not plain(op):
                    runs.append([op])
        if cur:
            runs.append(cur)
        groups = []
        for r in runs:
            m = _ceil(len(r), K)
            base, rem, pos = len(r) // m, len(r) % m, 0
            for g in range(m):                       # balanced split: 4 -> 2 + 2, not 3 + 1
                size = base + (1 if g < rem else 0)
                groups.append(r[pos:pos + size])
                pos += size
        # tile-order pass (see above), re-planning the layers whose direction changes
        if REVERSE_ORDER:
            dirs = {}
            for g in groups:
                for i, op in enumerate(g):
                    if not isinstance(op, ConvOp) or op.kernel != L.KERNEL_TC:
                        continue
                    d = (1 - dirs.get(id(op.io["srcs"][0]), 0)) if i == 0 else g[0].desc.reverse_items
                    if d != op.desc.reverse_items:
                        op.desc.reverse_items = d
                        op.rebuild()
                    for key in ("out", "out2", "dst2"):
                        if op.io.get(key) is not None:
                            dirs[id(op.io[key])] = d
        self.exec = [ChainOp(g, fused=True) if len(g) > 1 else g[0] for g in groups]

    def make_chain(self):
        """Fuse the whole launch list into one persistent launch when every op qualifies (ChainOp.chainable)."""
        if USE_CHAIN and self.prec.name == "bf16" and ChainOp.chainable(self.ops):
            try:
                self.chain = ChainOp(self.ops)
            except L.SrdError as e:          # a layer the chain planner does not take: one launch per layer instead
                self.chain, self.chain_error = None, str(e)
        return self.chain

    def count_kernels(self):
        """(launches per run, tcgen05 conv layers among the ops)"""
        n_tc = sum(1 for o in self.ops if isinstance(o, ConvOp) and o.kernel == L.KERNEL_TC)
        if self.chain is not None:
            return 2, n_tc                      # counter memset + the chain kernel
        return len(self.exec if self.exec is not None else self.ops), n_tc

    # ---- generic conv ----------------------------------------------------------------------
    def conv(self, pw, srcs, out=None, out_coff=0, out2=None, out2_coff=0, act=L.ACT_NONE, alpha=1.0,
             temb_off=None, temb_bstride=0, res1=None, res2=None, border=L.BORDER_KEEP, planar=None,
             planar_c=0, planar_clamp01=0, head=None, kernel=None, win8=False, lin=None, temb_static=None,
             dual=None):
        """srcs: list of (Act, c_off, cin); res1/res2: (Act, c_off, beta); head: (w[planar_c][N], b).
        win8: the single source is an 8-channel buffer read through SRD_SRC_WIN8 (3x3 conv as three K=64
        blocks).  lin: (Act, w[N][c], c) linear residual evaluated in the epilogue; temb_static: fp32 [N]
        vector added after the activation (e.g. the bias of the folded res_conv).  dual: (conv1x1 module, Act)
        second output of the same launch = conv1x1(srcs) + bias, no activation (ResnetBlock.res_conv)."""
        d = L.SrdConvDesc()
        d.dtype = self.prec.dtype
        d.kernel = self.prec.kernel if kernel is None else kernel
        a0 = srcs[0][0]
        d.n_src = len(srcs)
        d.B, d.H_in, d.W_in = a0.B, a0.H, a0.W
        for i, (a, coff, cin) in enumerate(srcs):
            assert (a.B, a.H, a.W) == (a0.B, a0.H, a0.W), "conv sources must share B/H/W"
            d.src[i].ptr, d.src[i].ct, d.src[i].c_off, d.src[i].cin = a.ptr, a.ct, coff, cin
        if win8:
            assert len(srcs) == 1 and a0.ct == 8 and not pw.transposed and pw.ksize == 3 and pw.stride == 1
            assert a0._flat.numel() - a0.t.numel() >= 56, "SRD_SRC_WIN8 source needs >= 112 B of slack"
            wd, wt, bias = pw.get_win8()
            d.src[0].cin, d.src[0].flags = 64, L.SRC_WIN8
            d.kernel = L.KERNEL_TC
        elif dual is not None:
            assert not pw.transposed and pw.ksize == 3 and pw.stride == 1
            wd, wt, bias, bias2 = pw.get_dual([s[2] for s in srcs], eff_weight(dual[0]), dual[0].bias)
            d.N2, d.bias2 = pw.N, bias2.data_ptr()
            d.dst2, d.dst2_ct, d.dst2_coff = dual[1].ptr, dual[1].ct, 0
            d.kernel = L.KERNEL_TC
        else:
            # 256-wide stride-1 convs on the tcgen05 path: two channel groups of 128 (see SrdConvDesc.n_split)
            n_split = 2 if (NSPLIT and self.prec.name == "bf16" and kernel is None and not pw.transposed and
                            pw.stride == 1 and pw.N == 256 and out is not None and out2 is None and planar is None and
                            head is None and lin is None and temb_bstride == 0 and act != L.ACT_LRELU02 and alpha == 1.0 and
                            a0.H >= 4 and out.ct % 8 == 0 and out_coff % 8 == 0 and
                            all(a.ct % 8 == 0 and coff % 8 == 0 for a, coff, _ in srcs) and
                            (res1 is None or (res1[0].ct % 8 == 0 and res1[1] % 8 == 0)) and
                            (res2 is None or (res2[0].ct % 8 == 0 and res2[1] % 8 == 0))) else 1
            d.n_split = n_split
            wd, wt, bias = pw.get([s[2] for s in srcs], n_split)
        if win8:
            d.stride, d.nky, d.nkx, d.ky0, d.kx0, d.up = 1, 3, 1, 0, 0, 1
        elif pw.transposed:
            S = pw.ksize // 2
            d.stride, d.nky, d.nkx, d.ky0, d.kx0, d.up = 1, 2, 2, 0, 0, S
        else:
            k = pw.ksize
            d.stride, d.nky, d.nkx, d.up = pw.stride, k, k, 1
            d.ky0 = d.kx0 = (3 - k) // 2
        d.N = pw.N
        d.w_direct, d.w_tc, d.bias = wd.data_ptr(), wt.data_ptr(), bias.data_ptr()
        d.act, d.alpha = act, alpha
        d.temb, d.temb_bstride = None, temb_bstride
        keep = [wd, wt, bias, [s[0] for s in srcs], out, out2, planar, dual, bias2 if dual is not None else None]
        if temb_static is not None:
            assert temb_off is None
            d.temb = temb_static.data_ptr()
            keep.append(temb_static)
        if lin is not None:
            d.lin_src, d.lin_ct, d.lin_c, d.lin_w = lin[0].ptr, lin[0].ct, lin[2], lin[1].data_ptr()
            d.kernel = L.KERNEL_TC
            keep.append(lin)
        if res1 is not None:
            d.res1, d.res1_ct, d.res1_coff, d.beta1 = res1[0].ptr, res1[0].ct, res1[1], res1[2]
            keep.append(res1[0])
        if res2 is not None:
            d.res2, d.res2_ct, d.res2_coff, d.beta2 = res2[0].ptr, res2[0].ct, res2[1], res2[2]
            keep.append(res2[0])
        d.H_out, d.W_out = a0.H * d.up // d.stride, a0.W * d.up // d.stride
        for o in (out, out2):
            if o is not None:
                assert (o.B, o.H, o.W) == (a0.B, d.H_out, d.W_out), "conv output buffer has the wrong shape"
        if out is not None:
            d.out, d.out_ct, d.out_coff = out.ptr, out.ct, out_coff
        if out2 is not None:
            d.out2, d.out2_ct, d.out2_coff = out2.ptr, out2.ct, out2_coff
        d.border = border
        # serpentine tile order: start where the producer of the first source finished (its output tail is still
        # in L2); REVERSE_ORDER = False keeps every layer front-to-back
        if REVERSE_ORDER and d.kernel != L.KERNEL_DIRECT:
            d.reverse_items = 1 - self._dir.get(id(a0), 0)        # unknown writers (converters, p_sample) go front to back
            for o in (out, out2, dual[1] if dual is not None else None):
                if o is not None:
                    self._dir[id(o)] = d.reverse_items
        if planar is not None:
            d.planar, d.planar_c, d.planar_clamp01 = planar.data_ptr(), planar_c, planar_clamp01
        if head is not None:
            d.head_w, d.head_b = head[0].data_ptr(), head[1].data_ptr()
            keep.append(head)
        io = {"srcs": [s_[0] for s_ in srcs], "res1": res1[0] if res1 is not None else None,
              "res2": res2[0] if res2 is not None else None, "lin": lin[0] if lin is not None else None,
              "out": out, "out2": out2, "dst2": dual[1] if dual is not None else None}
        op = ConvOp(d, keep, temb_off, io)
        self.ops.append(op)
        return op

    # ---- converters ------------------------------------------------------------------------
    def nchw_to_act(self, x_fn, act_buf, C_, scale=1.0, shift=0.0, border=L.BORDER_KEEP):
        """x_fn() -> fp32 NCHW tensor (resolved at run time so static buffers can be swapped)."""
        def f(ctx):
            x = x_fn()
            L.check(L.lib().srd_nchw_to_act(x.data_ptr(), act_buf.B, C_, act_buf.H, act_buf.W, scale, shift,
                                            act_buf.ptr, act_buf.ct, self.prec.dtype, border, _stream()),
                    "srd_nchw_to_act")
        self.ops.append(FnOp(f, (act_buf,)))

    def act_to_nchw(self, act_buf, c_off, C_, out):
        def f(ctx):
            L.check(L.lib().srd_act_to_nchw(act_buf.ptr, act_buf.ct, c_off, self.prec.dtype, act_buf.B, C_,
                                            act_buf.H, act_buf.W, out.data_ptr(), _stream()), "srd_act_to_nchw")
        self.ops.append(FnOp(f, (act_buf, out)))

    def groupnorm(self, src, C_, gn, out, act=L.ACT_MISH, temb_off=None, temb_bstride=0, res=None,
                  border=L.BORDER_KEEP):
        """out = act(GroupNorm(src)) + temb + res (`Block` :448-454 + ResnetBlock adds :474-480); gn: nn.GroupNorm."""
        assert (src.B, src.H, src.W) == (out.B, out.H, out.W) and src is not out
        gamma, beta = gn.weight.detach().float().contiguous(), gn.bias.detach().float().contiguous()
        stats = torch.empty(2 * src.B * gn.num_groups, dtype=torch.float32, device=self.device)
        a = L.SrdGroupNormArgs()
        a.src, a.src_ct, a.src_coff, a.dtype = src.ptr, src.ct, 0, self.prec.dtype
        a.out, a.out_ct, a.out_coff = out.ptr, out.ct, 0
        a.B, a.H, a.W, a.C, a.groups = src.B, src.H, src.W, C_, gn.num_groups
        a.act, a.border, a.eps, a.temb_bstride = act, border, gn.eps, temb_bstride
        a.gamma, a.beta, a.stats = gamma.data_ptr(), beta.data_ptr(), stats.data_ptr()
        if res is not None:
            a.res, a.res_ct, a.res_coff = res[0].ptr, res[0].ct, res[1]

        def f(ctx):
            a.temb = (ctx.temb_ptr + 4 * temb_off) if temb_off is not None else None
            L.check(L.lib().srd_groupnorm_act(C.byref(a), _stream()), "srd_groupnorm_act")
        self.ops.append(FnOp(f, (src, out, gamma, beta, stats, res)))

    def linear_attention(self, qkv, heads, dim_head, out):
        """`LinearAttention.forward` :514-522 between its 1x1 convs (k softmax over positions, k v^T, context^T q)."""
        def f(ctx):
            L.check(L.lib().srd_linear_attention(qkv.ptr, qkv.ct, 0, self.prec.dtype, qkv.B, qkv.H, qkv.W, heads,
                                                 dim_head, out.ptr, out.ct, 0, _stream()), "srd_linear_attention")
        self.ops.append(FnOp(f, (qkv, out)))

    def upsample2x(self, src, dst):
        def f(ctx):
            L.check(L.lib().srd_upsample_nearest2x(src.ptr, dst.ptr, self.prec.dtype, src.B, src.H, src.W,
                                                   src.ct, _stream()), "srd_upsample_nearest2x")
        assert src.ct == dst.ct
        self.ops.append(FnOp(f, (src, dst)))


# --------------------------------------------------------------------------------------------
# layer programs (mirror the reference's module graph; see model.py for the modules)
# --------------------------------------------------------------------------------------------
def M_hparams():
    from . import model as M
    return M.hparams


def eff_weight(module):
    """The weight a conv module computes with.  Under torch.nn.utils.weight_norm (`Unet.apply_weight_norm`
    :808-814) that is g * v / ||v|| (norm over every dim but 0): `module.weight` itself is only refreshed by
    nn.Conv2d's own forward pre-hook, which this package never runs, so it is folded here -- once per (g, v)
    version, the folded tensor is cached on the module."""
    g, v = getattr(module, "weight_g", None), getattr(module, "weight_v", None)
    if g is None or v is None:
        return module.weight
    sig = (g.data_ptr(), g._version, v.data_ptr(), v._version)
    hit = module.__dict__.get("_srd_wn")
    if hit is None or hit[0] != sig:
        w = torch._weight_norm(v.detach().float(), g.detach().float(), 0).contiguous()
        hit = (sig, w)
        module.__dict__["_srd_wn"] = hit
    return hit[1]


def _pc(module, attr="_srd_pack", **kw):
    """PackedConv attached to an nn.Conv2d / nn.ConvTranspose2d (cached on the module)."""
    p = module.__dict__.get(attr)
    w = eff_weight(module)                 # module.weight, or the folded weight under weight norm
    bias = module.bias
    if bias is None:                       # bias-free conv (LinearAttention.to_qkv :511): a cached zero vector
        bias = module.__dict__.get("_srd_zero_bias")
        if bias is None or bias.device != w.device:
            bias = torch.zeros(module.out_channels, dtype=torch.float32, device=w.device)
            module.__dict__["_srd_zero_bias"] = bias
    if p is None or p.weight is not w or p.bias_p is not bias:
        transposed = isinstance(module, torch.nn.ConvTranspose2d)
        stride = 1 if transposed else module.stride[0]
        p = PackedConv(w, bias, transposed=transposed, stride=stride)
        module.__dict__[attr] = p
    return p


def _pc_nobias(module):
    """PackedConv of an nn.Conv2d with a zero bias (for the linear part of a conv applied to a constant term)."""
    w = eff_weight(module)              # (under weight norm `module.weight` is a stale non-parameter attribute)
    z = module.__dict__.get("_srd_zero_bias2")
    if z is None or z.device != w.device:
        z = torch.zeros(module.out_channels, dtype=torch.float32, device=w.device)
        module.__dict__["_srd_zero_bias2"] = z
    p = module.__dict__.get("_srd_pack_nobias")
    if p is None or p.weight is not w or p.bias_p is not z:
        p = PackedConv(w, z, transposed=False, stride=module.stride[0])
        module.__dict__["_srd_pack_nobias"] = p
    return p


def build_resnet(bld, blk, srcs, dim_out, out, border, temb_off, temb_bstride=0, extra_res=None, mid_res=None):
    """ResnetBlock.forward :473-480.  srcs = K sources of the block input x (a list => fused concat).
    Returns `out` (Act with dim_out channels).  extra_res: (Act, c_off, beta) added at the tail
    (`x + cond` :830); mid_res: (Act, c_off, beta) added between block1 and block2 (`h += cond` :478)."""
    a0 = srcs[0][0]
    B, H, W = a0.B, a0.H, a0.W
    cin = sum(s[2] for s in srcs)
    h1 = bld.act(B, H, W, dim_out)
    use_temb = temb_off is not None
    if len(blk.block1.block) == 4:
        # GroupNorm variant (`Block` with groups > 0, :448-454): bias-only conv launches, each followed by one
        # normalise + Mish pass that also carries the adds of :476-480
        assert extra_res is None and mid_res is None
        t1, t2 = bld.act(B, H, W, dim_out), bld.act(B, H, W, dim_out)
        bld.conv(_pc(blk.block1.block[1]), srcs, out=t1, act=L.ACT_NONE, border=L.BORDER_KEEP)
        bld.groupnorm(t1, dim_out, blk.block1.block[2], h1, temb_off=temb_off,
                      temb_bstride=temb_bstride if use_temb else 0, border=L.BORDER_REFLECT)
        bld.conv(_pc(blk.block2.block[1]), [(h1, 0, dim_out)], out=t2, act=L.ACT_NONE, border=L.BORDER_KEEP)
        if cin != dim_out:
            r = bld.act(B, H, W, dim_out)
            bld.conv(_pc(blk.res_conv), srcs, out=r, act=L.ACT_NONE, border=L.BORDER_KEEP)
            res = (r, 0)
        else:
            assert len(srcs) == 1
            res = (srcs[0][0], srcs[0][1])
        bld.groupnorm(t2, dim_out, blk.block2.block[2], out, res=res, border=border)
        return out
    # Few-channel input (the Unet's first block, 3 -> 64) on the tcgen05 path: block1 reads the 8-channel
    # buffer through 8-pixel windows (K = 3 x 64 instead of 9 x 64), and the 1x1 res_conv (:471,480)
    # is evaluated inside block2's epilogue (3 FMAs per output) instead of a launch of its own.
    packed = (bld.prec.name == "bf16" and len(srcs) == 1 and cin <= 4 and a0.ct == 8 and srcs[0][1] == 0 and
              cin != dim_out and dim_out % 64 == 0 and W % 8 == 0 and (temb_bstride == 0 or not use_temb) and
              extra_res is None and mid_res is None and a0._flat.numel() - a0.t.numel() >= 56)
    if packed:
        bld.conv(_pc(blk.block1.block[1]), srcs, out=h1, act=L.ACT_MISH, temb_off=temb_off, border=L.BORDER_REFLECT,
                 win8=True)
        lw = eff_weight(blk.res_conv).detach().float().reshape(dim_out, cin).contiguous()
        lb = blk.res_conv.bias.detach().float().contiguous()
        bld.conv(_pc(blk.block2.block[1]), [(h1, 0, dim_out)], out=out, act=L.ACT_MISH, lin=(a0, lw, cin),
                 temb_static=lb, border=border)
        return out
    # channel-changing block on the tcgen05 path: the 1x1 res_conv is a second output of block1's launch
    # (same activation tiles; SrdConvDesc.N2) when the doubled accumulator still leaves two TMEM stages
    fused_res = (cin != dim_out and FUSE_RES_CONV and mid_res is None and bld.prec.name == "bf16" and dim_out % 64 == 0 and dim_out <= FUSE_RES_MAX_N and H >= 4 and
                 (temb_bstride == 0 or not use_temb) and all(s[0].ct % 8 == 0 and s[1] % 8 == 0 for s in srcs))
    r = bld.act(B, H, W, dim_out) if cin != dim_out else None
    bld.conv(_pc(blk.block1.block[1]), srcs, out=h1, act=L.ACT_MISH, temb_off=temb_off,
             temb_bstride=temb_bstride if use_temb else 0, border=L.BORDER_REFLECT, res1=mid_res,
             dual=(blk.res_conv, r) if fused_res else None)
    if cin != dim_out:
        if not fused_res:
            bld.conv(_pc(blk.res_conv), srcs, out=r, act=L.ACT_NONE, border=L.BORDER_KEEP)
        res1 = (r, 0, 1.0)
    else:
        assert len(srcs) == 1
        res1 = (srcs[0][0], srcs[0][1], 1.0)
    bld.conv(_pc(blk.block2.block[1]), [(h1, 0, dim_out)], out=out, act=L.ACT_MISH, res1=res1,
             res2=extra_res, border=border)
    return out


def build_rdb_input(bld, x, nf, gc):
    """(nf + 4 gc)-channel dense-block buffer whose first nf channels hold the fp32 NCHW tensor `x` (zero ring)."""
    B, _, H, W = x.shape
    cat = bld.act(B, H, W, nf + 4 * gc)
    bld.nchw_to_act(lambda: x, cat, nf)
    return cat


def build_rdb(bld, rdb, cat, nf, gc, out, alpha, res1, res2=None, out2=None):
    """ResidualDenseBlock_5C.forward :685-690 on a dense-block buffer `cat` whose channels [0, nf) hold x:
    conv1..4 append their gc output channels in place (the `torch.cat`s), conv5 writes
    alpha * conv5(cat) + beta1 * res1 [+ beta2 * res2] into channels [0, nf) of `out` (and `out2`)."""
    K, LR = L.BORDER_KEEP, L.ACT_LRELU02
    for k, conv in enumerate((rdb.conv1, rdb.conv2, rdb.conv3, rdb.conv4)):
        bld.conv(_pc(conv), [(cat, 0, nf + k * gc)], out=cat, out_coff=nf + k * gc, act=LR, border=K)
    bld.conv(_pc(rdb.conv5), [(cat, 0, nf + 4 * gc)], out=out, out2=out2, alpha=alpha, res1=res1, res2=res2, border=K)


class UnetProgram:
    """Unet.forward :820-847 for fixed (B, H, W) as a static launch list.

    Inputs: `xin` (padded NHWC copy of x_t, 3 real channels, reflect ring) and `cond`
    (cond_proj output, 64 channels).  Output: `eps` fp32 NCHW.  The time embedding comes from a
    precomputed table row addressed through RunCtx.temb_ptr.
    """

    def __init__(self, unet, B, H, W, precision, device, cond_act=None, per_sample_temb=False, dcond_act=None):
        from . import model as M
        hp = M.hparams
        bld = self.bld = Builder(device, precision)
        self.B, self.H, self.W = B, H, W
        dims = unet._dims
        nlev = len(dims) - 1
        if H % (2 ** (nlev - 1)) or W % (2 ** (nlev - 1)):
            raise ValueError("Unet input %dx%d must be divisible by %d" % (H, W, 2 ** (nlev - 1)))
        self.xin = bld.act(B, H, W, 3, slack=64)       # slack: read through SRD_SRC_WIN8 windows
        self.cond = cond_act if cond_act is not None else bld.act(B, H, W, dims[1])
        self.eps = torch.empty((B, unet._out_dim, H, W), dtype=torch.float32, device=device)
        # time-table offsets, in the order of unet.resnet_blocks()
        offs, o = {}, 0
        for blk in unet.resnet_blocks():
            offs[id(blk)] = o
            o += blk._dim_out
        self.temb_S = o
        temb_bstride = o if per_sample_temb else 0        # 0: one table row shared by the whole batch
        RB, RK = L.BORDER_REFLECT, L.BORDER_KEEP
        x_src = [(self.xin, 0, dims[0])]
        hs = []
        h, w = H, W
        for i in range(nlev):
            r0, r1, down = unet.downs[i]
            d = dims[i + 1]
            x1 = bld.act(B, h, w, d)
            build_resnet(bld, r0, x_src, d, x1, RB, offs[id(r0)], temb_bstride)
            x2 = bld.act(B, h, w, d)
            extra = (self.cond, 0, 1.0) if (i == 0 and dcond_act is None) else None          # `x = x + cond` :830
            build_resnet(bld, r1, [(x1, 0, d)], d, x2, RB, offs[id(r1)], temb_bstride, extra_res=extra)
            # (hparams['up_input']: `+ up_proj(img_lr_up)` :831-832 is already part of `cond`, see CondProgram)
            hs.append((x2, d))
            if i < nlev - 1:
                xd = bld.act(B, h // 2, w // 2, d)
                bld.conv(_pc(down.conv[1]), [(x2, 0, d)], out=xd, act=L.ACT_NONE, border=RB,
                         res2=(dcond_act, 0, 1.0) if (i == 0 and dcond_act is not None) else None)   # :494-503
                x_src = [(xd, 0, d)]
                h, w = h // 2, w // 2
            else:
                x_src = [(x2, 0, d)]
        d = dims[-1]
        m1 = bld.act(B, h, w, d)
        build_resnet(bld, unet.mid_block1, x_src, d, m1, RB, offs[id(unet.mid_block1)], temb_bstride)
        if hp["use_attn"]:
            # mid_attn = Residual(Rezero(LinearAttention)) :788-789,836-837: x + g * to_out(attn(to_qkv(x)))
            att = unet.mid_attn.fn.fn
            hidden = att.to_out.in_channels
            qkv, ao, m1a = bld.act(B, h, w, 3 * hidden), bld.act(B, h, w, hidden), bld.act(B, h, w, d)
            bld.conv(_pc(att.to_qkv), [(m1, 0, d)], out=qkv, border=RK)
            bld.linear_attention(qkv, att.heads, hidden // att.heads, ao)
            bld.conv(_pc(att.to_out), [(ao, 0, hidden)], out=m1a, alpha=float(unet.mid_attn.fn.g.detach().float().item()),
                     res1=(m1, 0, 1.0), border=RB)
            m1 = m1a
        m2 = bld.act(B, h, w, d)
        build_resnet(bld, unet.mid_block2, [(m1, 0, d)], d, m2, RB, offs[id(unet.mid_block2)], temb_bstride)
        x, xd_ = m2, d
        for j in range(nlev - 1):
            r0, r1, up = unet.ups[j]
            skip, sd = hs.pop()
            dout = dims[nlev - 1 - j]
            u1 = bld.act(B, h, w, dout)
            build_resnet(bld, r0, [(x, 0, xd_), (skip, 0, sd)], dout, u1, RB, offs[id(r0)], temb_bstride)   # cat :842
            u2 = bld.act(B, h, w, dout)
            build_resnet(bld, r1, [(u1, 0, dout)], dout, u2, RK, offs[id(r1)], temb_bstride)   # feeds convT: zero ring
            h, w = h * 2, w * 2
            u3 = bld.act(B, h, w, dout)
            bld.conv(_pc(up.conv[0]), [(u2, 0, dout)], out=u3, act=L.ACT_NONE, border=RB)   # Upsample :483-491
            x, xd_ = u3, dout
        fc0, fc1 = unet.final_conv[0].block[1], unet.final_conv[1]
        od = unet._out_dim
        if bld.prec.name == "bf16" and od <= 4:
            hw = eff_weight(fc1).detach().float().reshape(od, xd_).contiguous()
            hb = fc1.bias.detach().float().contiguous()
            bld.conv(_pc(fc0), [(x, 0, xd_)], out=None, act=L.ACT_MISH, planar=self.eps, planar_c=od,
                     head=(hw, hb))
        else:
            f = bld.act(B, h, w, xd_)
            bld.conv(_pc(fc0), [(x, 0, xd_)], out=f, act=L.ACT_MISH, border=RK)
            bld.conv(_pc(fc1), [(f, 0, xd_)], planar=self.eps, planar_c=od)
        if not per_sample_temb:
            bld.make_chain()
            if bld.chain is None:
                bld.fuse_segments()

    def run(self, temb_ptr):
        self.bld.run(RunCtx(temb_ptr))


class RRDBProgram:
    """RRDBNet.forward(x, get_fea=True) :727-749 for fixed (B, h, w)."""

    def __init__(self, rrdb, B, h, w, precision, device, want_all_feas=True):
        from . import model as M
        hp = M.hparams
        bld = self.bld = Builder(device, precision)
        nf, gc, nb = rrdb._nf, rrdb._gc, len(rrdb.RRDB_trunk)
        self.nf, self.nb, self.B, self.h, self.w = nf, nb, B, h, w
        K, LR = L.BORDER_KEEP, L.ACT_LRELU02
        self.xin = bld.act(B, h, w, rrdb._in_nc)
        ctot = nf + 4 * gc
        cats = [bld.act(B, h, w, ctot) for _ in range(2)]
        self.fea_first = bld.act(B, h, w, nf)
        self.feas = [bld.act(B, h, w, nf) for _ in range(nb + 1)]
        # conv_first :730 (input already (x+1)/2, done by the converter)
        bld.conv(_pc(rrdb.conv_first), [(self.xin, 0, rrdb._in_nc)], out=cats[0], out2=self.fea_first, border=K)
        cur = 0
        for i, blk in enumerate(rrdb.RRDB_trunk):
            x_rrdb = self.fea_first if i == 0 else self.feas[i - 1]
            for j, rdb in enumerate((blk.RDB1, blk.RDB2, blk.RDB3)):
                c, n = cats[cur], cats[cur ^ 1]
                if j < 2:       # x5*0.2 + x :690
                    build_rdb(bld, rdb, c, nf, gc, n, alpha=0.2, res1=(c, 0, 1.0))
                else:           # (x5*0.2 + x)*0.2 + x_rrdb :690,706
                    build_rdb(bld, rdb, c, nf, gc, n, alpha=0.04, res1=(c, 0, 0.2), res2=(x_rrdb, 0, 1.0),
                              out2=self.feas[i])
                cur ^= 1
        # trunk_conv + fea_first :734-736
        bld.conv(_pc(rrdb.trunk_conv), [(cats[cur], 0, nf)], out=self.feas[nb], res1=(self.fea_first, 0, 1.0),
                 border=K)
        # HR head :738-745
        self.out = torch.empty((B, rrdb._out_nc, h * hp["sr_scale"], w * hp["sr_scale"]), dtype=torch.float32,
                               device=device)
        fea, hh, ww = self.feas[nb], h, w
        ups = [rrdb.upconv1, rrdb.upconv2] + ([rrdb.upconv3] if hp["sr_scale"] == 8 else [])
        for conv in ups:
            u = bld.act(B, hh * 2, ww * 2, nf)
            bld.upsample2x(fea, u)
            hh, ww = hh * 2, ww * 2
            v = bld.act(B, hh, ww, nf)
            bld.conv(_pc(conv), [(u, 0, nf)], out=v, act=LR, border=K)
            fea = v
        hr = bld.act(B, hh, ww, nf)
        bld.conv(_pc(rrdb.HRconv), [(fea, 0, nf)], out=hr, act=LR, border=K)      # lrelu(fea_hr) :743
        bld.conv(_pc(rrdb.conv_last), [(hr, 0, nf)], planar=self.out, planar_c=rrdb._out_nc, planar_clamp01=1)

    def load_input(self, img_lr):
        L.check(L.lib().srd_nchw_to_act(img_lr.data_ptr(), self.B, img_lr.shape[1], self.h, self.w, 0.5, 0.5,
                                        self.xin.ptr, self.xin.ct, self.bld.prec.dtype, L.BORDER_KEEP,
                                        _stream()), "srd_nchw_to_act")   # (x+1)/2 :729

    def run(self):
        self.bld.run()

    def feas_nchw(self):
        outs = []
        for a in self.feas:
            o = torch.empty((self.B, self.nf, self.h, self.w), dtype=torch.float32, device=a.t.device)
            L.check(L.lib().srd_act_to_nchw(a.ptr, a.ct, 0, self.bld.prec.dtype, self.B, self.nf, self.h, self.w,
                                            o.data_ptr(), _stream()), "srd_act_to_nchw")
            outs.append(o)
        return outs


class CondProgram:
    """cond = cond_proj(cat(cond[2::3], 1)) :825 -- ConvTranspose2d(3*cond_dim, dim, 2S, S, S/2),
    time-invariant, so computed once per image."""

    def __init__(self, unet, fea_acts, precision, device):
        bld = self.bld = Builder(device, precision)
        sel = fea_acts[2::3]
        a0 = sel[0]
        S = unet.cond_proj.stride[0]
        cin_total = unet.cond_proj.weight.shape[0]
        cins = [unet._cond_dim] * len(sel)
        if sum(cins) != cin_total:
            raise RuntimeError(
                "cond_proj expects %d input channels but the RRDB features provide %d (%d maps of %d): "
                "cond_dim / rrdb_num_feat / rrdb_num_block mismatch" % (cin_total, sum(cins), len(sel), cins[0]))
        # `x = x + cond` :830 sits right before the first Downsample (a ReflectionPad2d + stride-2 conv, both linear)
        # and that sum has no other reader (the level-0 skip is never popped), so when FOLD_COND the
        # Unet adds `Downsample_linear(cond)` -- time-invariant, a quarter of the pixels -- to the Downsample's
        # output instead of reading the full-resolution cond tensor in every step.
        down = unet.downs[0][2]
        self.fold = FOLD_COND and len(unet.downs) > 1 and not isinstance(down, torch.nn.Identity)
        ring = L.BORDER_REFLECT if self.fold else L.BORDER_KEEP
        up_in = bool(M_hparams()['res'] and M_hparams()['up_input'])
        self.cond = bld.act(a0.B, a0.H * S, a0.W * S, unet._dims[1])
        bld.conv(_pc(unet.cond_proj), [(a, 0, c) for a, c in zip(sel, cins)], out=self.cond,
                 border=L.BORDER_KEEP if up_in else ring)
        # hparams['up_input'] (:803-806, 831-832): x = x + cond + up_proj(img_lr_up) -- time-invariant as well, so
        # it is folded into the cached cond tensor.  `lr_up` (3 channels, reflect ring) is filled by the caller.
        self.lr_up = None
        if up_in:
            self.lr_up = bld.act(a0.B, a0.H * S, a0.W * S, 3)
            raw, self.cond = self.cond, bld.act(a0.B, a0.H * S, a0.W * S, unet._dims[1])
            bld.conv(_pc(unet.up_proj[1]), [(self.lr_up, 0, 3)], out=self.cond, res1=(raw, 0, 1.0), border=ring)
        self.dcond = None
        if self.fold:
            c = self.cond
            self.dcond = bld.act(c.B, c.H // 2, c.W // 2, unet._dims[1])
            bld.conv(_pc_nobias(down.conv[1]), [(c, 0, unet._dims[1])], out=self.dcond, border=L.BORDER_KEEP)

    def load_lr_up(self, img_lr_up):
        """img_lr_up: fp32 NCHW [B, 3, H, W] (only when the Unet has `up_proj`)."""
        a = self.lr_up
        L.check(L.lib().srd_nchw_to_act(img_lr_up.data_ptr(), a.B, 3, a.H, a.W, 1.0, 0.0, a.ptr, a.ct,
                                        self.bld.prec.dtype, L.BORDER_REFLECT, _stream()), "srd_nchw_to_act")

    def run(self):
        self.bld.run()


# --------------------------------------------------------------------------------------------
# time table
# --------------------------------------------------------------------------------------------
def time_table(unet, t_values):
    """[n_t][S] fp32 table of every ResnetBlock's `mlp(time_emb)` :821-822,:476 for the given t's."""
    dev = unet.mlp[0].weight.device
    dim = unet._dim
    half = dim // 2
    e = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half) * -e).to(dev).float().contiguous()     # :415-416 (CPU, then moved)
    blocks = unet.resnet_blocks()
    w_blk = torch.cat([b.mlp[1].weight.detach().float() for b in blocks], 0).contiguous()
    b_blk = torch.cat([b.mlp[1].bias.detach().float() for b in blocks], 0).contiguous()
    S = w_blk.shape[0]
    t_values = t_values.to(dev).float().contiguous()
    out = torch.empty((t_values.numel(), S), dtype=torch.float32, device=dev)
    w0, b0 = unet.mlp[0].weight.detach().float().contiguous(), unet.mlp[0].bias.detach().float().contiguous()
    w2, b2 = unet.mlp[2].weight.detach().float().contiguous(), unet.mlp[2].bias.detach().float().contiguous()
    L.check(L.lib().srd_time_table(t_values.data_ptr(), t_values.numel(), dim, freqs.data_ptr(), w0.data_ptr(),
                                   b0.data_ptr(), w2.data_ptr(), b2.data_ptr(), w_blk.data_ptr(), b_blk.data_ptr(),
                                   S, out.data_ptr(), _stream()), "srd_time_table")
    return out


def plan_signature(diffusion):
    """Everything a cached SamplerPlan (and its captured graph) depends on besides (B, h, w, precision, device)."""
    from . import model as M
    hp = M.hparams
    return (params_signature(diffusion), diffusion.num_timesteps, hp["res_rescale"], hp["clip_input"], bool(hp["res"]),
            bool(hp.get("up_input", False)), bool(hp.get("use_attn", False)), hp["sr_scale"])


def params_signature(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


# --------------------------------------------------------------------------------------------
# sampler plan: GaussianDiffusion.sample :1062-1091 as (optionally) ONE CUDA graph
# --------------------------------------------------------------------------------------------
COEF_NAMES = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
              "posterior_mean_coef2", "posterior_log_variance_clipped")


def coef_table(diffusion):
    """[T][5] fp32 rows consumed by srd_p_sample_step (the six `extract` gathers of :968-981)."""
    return torch.stack([getattr(diffusion, n).detach().float() for n in COEF_NAMES], 1).contiguous()


def p_sample_step(x, eps, noise, coef, t_i32, x_out, clip_denoised=True, seed=0, step=0, image_offset=0,
                  x0_out=None, img_out=None, img_lr_up=None, res_rescale=1.0, clip_input=True, nhwc_out=None,
                  act_dtype=L.BF16, seed_dev=None):
    a = L.SrdPSampleArgs()
    B, C_, H, W = x.shape
    a.x, a.eps, a.noise = x.data_ptr(), eps.data_ptr(), (noise.data_ptr() if noise is not None else None)
    a.coef, a.t = coef.data_ptr(), t_i32.data_ptr()
    a.B, a.C, a.H, a.W = B, C_, H, W
    a.clip_denoised = 1 if clip_denoised else 0
    a.seed, a.step, a.image_offset = seed, step, image_offset
    a.x_out = x_out.data_ptr()
    a.x0_out = x0_out.data_ptr() if x0_out is not None else None
    if img_out is not None:
        a.img_out, a.img_lr_up = img_out.data_ptr(), img_lr_up.data_ptr()
    a.res_rescale, a.clip_input = res_rescale, 1 if clip_input else 0
    if nhwc_out is not None:
        a.nhwc_out, a.nhwc_ct = nhwc_out.ptr, nhwc_out.ct
    a.act_dtype = act_dtype
    a.seed_ptr = seed_dev.data_ptr() if seed_dev is not None else None
    L.check(L.lib().srd_p_sample_step(C.byref(a), _stream()), "srd_p_sample_step")


class SamplerPlan:
    """Everything `sample()` needs for one (B, h, w, T, precision): static buffers, the RRDB /
    cond_proj / Unet launch lists, the time and coefficient tables, and the captured graph."""

    def __init__(self, diffusion, B, h, w, precision, device):
        from . import model as M
        hp = M.hparams
        self.prec = Precision(precision)
        self.B, self.h, self.w = B, h, w
        s = hp["sr_scale"]
        self.H, self.W = h * s, w * s
        self.T = diffusion.num_timesteps
        unet, rrdb = diffusion.denoise_fn, diffusion.rrdb
        self.rrdb_prog = RRDBProgram(rrdb, B, h, w, self.prec, device)
        self.cond_prog = CondProgram(unet, self.rrdb_prog.feas, self.prec, device)
        self.unet_prog = UnetProgram(unet, B, self.H, self.W, self.prec, device, cond_act=self.cond_prog.cond,
                                     dcond_act=self.cond_prog.dcond)
        C_ = unet._out_dim
        f32 = dict(dtype=torch.float32, device=device)
        self.x = torch.empty((B, C_, self.H, self.W), **f32)
        self.img_lr = torch.empty((B, rrdb._in_nc, h, w), **f32)
        self.img_lr_up = torch.empty((B, C_, self.H, self.W), **f32)
        self.img = torch.empty((B, C_, self.H, self.W), **f32)
        self.coef = coef_table(diffusion)
        self.t_idx = torch.arange(self.T, dtype=torch.int32, device=device)[:, None].repeat(1, B).contiguous()
        self.temb = time_table(unet, torch.arange(self.T, dtype=torch.float32))
        self.res_rescale = float(hp["res_rescale"])
        self.clip_input = bool(hp["clip_input"])
        self.noise_buf = None
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=device)   # Philox seed, read by the kernels
        self.graphs = {}
        self.sig = plan_signature(diffusion)

    # ---- pieces -----------------------------------------------------------------------------
    def encode(self):
        """RRDB encoder + cond_proj: once per image (cached across all T steps)."""
        self.rrdb_prog.load_input(self.img_lr)
        self.rrdb_prog.run()
        if self.cond_prog.lr_up is not None:
            self.cond_prog.load_lr_up(self.img_lr_up)
        self.cond_prog.run()

    def load_x(self):
        L.check(L.lib().srd_nchw_to_act(self.x.data_ptr(), self.B, self.x.shape[1], self.H, self.W, 1.0, 0.0,
                                        self.unet_prog.xin.ptr, self.unet_prog.xin.ct, self.prec.dtype,
                                        L.BORDER_REFLECT, _stream()), "srd_nchw_to_act")

    def step(self, i, k, noise, seed, image_offset, x0_out=None, fuse_img=True):
        """One p_sample :1051-1060 at timestep i (k-th step of the loop)."""
        self.unet_prog.run(self.temb.data_ptr() + 4 * i * self.temb.shape[1])
        last = (i == 0)
        p_sample_step(self.x, self.unet_prog.eps, noise, self.coef, self.t_idx[i], self.x, True, seed, k + 1,
                      image_offset, x0_out=x0_out,
                      img_out=self.img if (last and fuse_img) else None, img_lr_up=self.img_lr_up,
                      res_rescale=self.res_rescale, clip_input=self.clip_input,
                      nhwc_out=None if last else self.unet_prog.xin, act_dtype=self.prec.dtype,
                      seed_dev=self.seed_dev)

    def loop(self, noises, seed, image_offset):
        self.load_x()
        for k, i in enumerate(reversed(range(self.T))):
            self.step(i, k, None if noises is None else noises[k], seed, image_offset)

    def ensure_noise_buf(self):
        if self.noise_buf is None:
            self.noise_buf = torch.empty((self.T,) + tuple(self.x.shape), dtype=torch.float32, device=self.x.device)
        return self.noise_buf

    def noise_bytes(self):
        return 4 * self.T * self.x.numel()

    def run_graph(self, use_noise_buf, seed, image_offset):
        """Replay (capturing on first use) encode + the whole T-step loop as one CUDA graph.
        The Philox seed lives in device memory (`seed_dev`), so a new seed does not need a new graph;
        the shard's image offset is baked into the kernel arguments and keys the graph."""
        self.seed_dev.fill_(seed & 0x7FFFFFFFFFFFFFFF)
        key = (use_noise_buf, image_offset if not use_noise_buf else 0)
        g = self.graphs.pop(key, None)
        if g is None:
            from . import model as M
            while len(self.graphs) >= max(1, M.runtime.graph_cache_size):     # LRU (dicts keep insertion order)
                self.graphs.pop(next(iter(self.graphs)))
            noises = self.ensure_noise_buf() if use_noise_buf else None
            # warm-up outside capture (lazy module/func loading, cudaFuncSetAttribute)
            x_save = self.x.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.encode()
                self.load_x()
                self.step(self.T - 1, 0, None if noises is None else noises[0], seed, image_offset)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.x.copy_(x_save)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.encode()
                self.loop(noises, seed, image_offset)
        self.graphs[key] = g                                                  # most recently used last
        g.replay()

    def n_launches(self):
        n_r, _ = self.rrdb_prog.bld.count_kernels()
        n_c, _ = self.cond_prog.bld.count_kernels()
        n_u, _ = self.unet_prog.bld.count_kernels()
        return (n_r + 1) + n_c + 1 + self.T * (n_u + 1)
