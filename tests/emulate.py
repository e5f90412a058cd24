"""CPU emulation of the convolution *descriptor semantics* (include/srdiff_b200.h) in plain torch:
padded-NHWC sources, tap windows, sub-pixel phases, and both packed weight layouts.  It checks the
host logic (weight packing order, tap tables, ring conventions) against F.conv2d /
F.conv_transpose2d without a GPU.  The device kernels index exactly like this."""
import torch


def to_padded_nhwc(x, reflect):
    """[B,C,H,W] -> [B,H+2,W+2,C] with a reflected or zero ring."""
    mode = "reflect" if reflect else "constant"
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1), mode=mode)
    return xp.permute(0, 2, 3, 1).contiguous()


def emulate_direct(srcs, cins, w_direct, N, stride, nk, k0, up):
    """srcs: padded NHWC tensors; w_direct [P][K][N] with K order (src, ky, kx, c)."""
    B, Hp, Wp, _ = srcs[0].shape
    H, W = Hp - 2, Wp - 2
    Hm, Wm = H // stride, W // stride
    out = torch.zeros((B, Hm * up, Wm * up, N))
    for ph in range(up * up):
        r, s = ph // up, ph % up
        ky0 = kx0 = k0
        if up > 1:
            ky0, kx0 = (1 if r >= up // 2 else 0), (1 if s >= up // 2 else 0)
        cols = []
        for src, cin in zip(srcs, cins):
            for ky in range(nk):
                for kx in range(nk):
                    ys = torch.arange(Hm) * stride + ky0 + ky
                    xs = torch.arange(Wm) * stride + kx0 + kx
                    cols.append(src[:, ys][:, :, xs][..., :cin])
        A = torch.cat(cols, -1)                       # [B,Hm,Wm,K]
        out[:, r::up, s::up] = A @ w_direct[ph]
    return out


def emulate_tc(srcs, cins, w_tc, N, stride, nk, k0, up):
    """w_tc [P][KB][Npad][64] bf16 with K-block order (src, chunk, kx, ky*); stride 2: per (src, chunk) the nine taps in
    the order of the (row parity, column parity) halo boxes (include/srdiff_b200.h, SrdConvDesc.w_tc), or -- with
    SRD_S2_HALO=0 -- (kx, ky*) with ky* = [0,2,1]."""
    from dgm_img_super_resolution_b200 import _lib
    B, Hp, Wp, _ = srcs[0].shape
    H, W = Hp - 2, Wp - 2
    Hm, Wm = H // stride, W // stride
    out = torch.zeros((B, Hm * up, Wm * up, N))
    w = w_tc.float()
    if w.dim() == 3:
        w = w[None]
    ky_order = [0, 2, 1] if stride == 2 else list(range(nk))
    for ph in range(up * up):
        r, s = ph // up, ph % up
        ky0 = kx0 = k0
        if up > 1:
            ky0, kx0 = (1 if r >= up // 2 else 0), (1 if s >= up // 2 else 0)
        kb = 0
        acc = torch.zeros((B, Hm, Wm, w.shape[2]))
        for src, cin in zip(srcs, cins):
            for ch in range((cin + 63) // 64):
                c0, c1 = ch * 64, min(cin, ch * 64 + 64)
                taps = [(ky, kx) for kx in range(nk) for ky in ky_order]
                if stride == 2 and _lib.lib().srd_tc_s2_halo():
                    taps = [(0, 0), (0, 2), (2, 0), (2, 2), (0, 1), (2, 1), (1, 0), (1, 2), (1, 1)]
                for ky, kx in taps:
                    ys = torch.arange(Hm) * stride + ky0 + ky
                    xs = torch.arange(Wm) * stride + kx0 + kx
                    a = torch.zeros((B, Hm, Wm, 64))
                    a[..., :c1 - c0] = src[:, ys][:, :, xs][..., c0:c1]     # TMA zero-fills c >= cin
                    acc += a @ w[ph, kb].t()
                    kb += 1
        assert kb == w.shape[1]
        out[:, r::up, s::up] = acc[..., :N]
    return out
