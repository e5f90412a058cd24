"""the reference's quality metrics on the GPU (srdiff/model.py:146-222 `SSIM` / `ssim`, and the
PSNR of `Measure` :1499-1542), behind the same names and call signatures.  CUDA tensors only; the arithmetic is
`srd_ssim`, `srd_f32_to_u8_round`, `srd_sq_err_u8` of libsrdiff_b200.so (csrc/variant_ops.cu).

Also here: the MATLAB-compatible `imresize` (:1162-1306; host-side `contributions`, one CUDA pass per dimension) and
`Measure.ssim` (scikit-image's uniform-window SSIM; `srd_ssim_uniform_u8`).
Not built: `Measure.lpips` -- it needs the pretrained LPIPS/AlexNet weights, which neither the reference nor this
image ships; `Measure.measure` therefore reports psnr / ssim / lr_psnr and adds 'lpips' only when the caller hands
the constructor a callable for it.
"""
import math

import numpy as np
import torch

from . import _lib as L
from . import engine as E


def gaussian(window_size, sigma):
    """:149-151 (the same fp32 normalisation as the reference)."""
    gauss = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def _ssim_per_image(img1, img2, window_size, pre_scale, pre_shift):
    E._require_cuda(img1, "img1")
    E._require_cuda(img2, "img2")
    if img1.shape != img2.shape or img1.dim() != 4:
        raise RuntimeError("ssim expects two [B, C, H, W] tensors of the same shape")
    a, b = img1.contiguous().float(), img2.contiguous().float()
    B, C_, H, W = a.shape
    win = gaussian(window_size, 1.5).to(a.device).contiguous()
    ws = torch.empty(L.lib().srd_ssim_workspace(B, C_, H, W), dtype=torch.float32, device=a.device)
    out = torch.empty(B, dtype=torch.float32, device=a.device)
    L.check(L.lib().srd_ssim(a.data_ptr(), b.data_ptr(), B, C_, H, W, window_size, win.data_ptr(), pre_scale,
                             pre_shift, ws.data_ptr(), out.data_ptr(), E._stream()), "srd_ssim")
    return out


def ssim(img1, img2, window_size=11, size_average=True):
    """:212-221 -- functional form, images used as given."""
    v = _ssim_per_image(img1, img2, window_size, 1.0, 0.0)
    return v.mean() if size_average else v


class SSIM(torch.nn.Module):
    """:185-209 -- images in [-1, 1] are mapped to [0, 1] first (`img * 0.5 + 0.5`)."""

    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average

    def forward(self, img1, img2):
        v = _ssim_per_image(img1, img2, self.window_size, 0.5, 0.5)
        return v.mean() if self.size_average else v


def to_uint8_round(img):
    """`np.round((img + 1) * 127.5).clip(0, 255).astype(np.uint8)` (:1516-1518), layout unchanged."""
    E._require_cuda(img, "img")
    x = img.contiguous().float()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    L.check(L.lib().srd_f32_to_u8_round(x.data_ptr(), x.numel(), out.data_ptr(), E._stream()), "srd_f32_to_u8_round")
    return out


def psnr_u8(a, b, data_range=255):
    """`skimage.metrics.peak_signal_noise_ratio(a, b, data_range=255)` on uint8 CUDA tensors:
    10 * log10(data_range^2 / mean((a - b)^2)), the squared-error sum accumulated exactly in 64-bit integers."""
    E._require_cuda(a, "a")
    if a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.shape != b.shape:
        raise RuntimeError("psnr_u8 expects two uint8 tensors of the same shape")
    a, b = a.contiguous(), b.contiguous()
    acc = torch.zeros(1, dtype=torch.int64, device=a.device)
    L.check(L.lib().srd_sq_err_u8(a.data_ptr(), b.data_ptr(), a.numel(), acc.data_ptr(), E._stream()), "srd_sq_err_u8")
    err = int(acc.item()) / float(a.numel())
    return float("inf") if err == 0 else 10.0 * math.log10((data_range ** 2) / err)


# ------------------------------------------------------------------------------------------ imresize
def _cubic(x):
    """:1188-1195 -- Keys cubic, a = -0.5."""
    ax = np.abs(np.asarray(x, dtype=np.float64))
    ax2, ax3 = ax * ax, ax * ax * ax
    return (1.5 * ax3 - 2.5 * ax2 + 1) * (ax <= 1) + (-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2) * ((1 < ax) & (ax <= 2))


def _triangle(x):
    """:1179-1185"""
    x = np.asarray(x, dtype=np.float64)
    return (x + 1) * ((x >= -1) & (x < 0)) + (1 - x) * ((x <= 1) & (x >= 0))


def contributions(in_length, out_length, scale, kernel, k_width):
    """:1198-1218 -- per output sample the taps (float64 weights, mirrored source indices) of the, when shrinking
    widened, interpolation kernel; all-zero tap columns dropped.  Returns (weights [out][P], indices [out][P])."""
    if scale < 1:
        def h(x):
            return scale * kernel(scale * x)
        kernel_width = 1.0 * k_width / scale
    else:
        h, kernel_width = kernel, k_width
    u = np.arange(1, out_length + 1, dtype=np.float64) / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kernel_width / 2)
    P = int(math.ceil(kernel_width)) + 2
    ind = (left[:, None] + np.arange(P) - 1).astype(np.int32)
    w = h(u[:, None] - ind - 1)
    w = w / w.sum(axis=1, keepdims=True)
    mirror = np.concatenate((np.arange(in_length), np.arange(in_length - 1, -1, -1))).astype(np.int32)
    ind = mirror[np.mod(ind, mirror.size)]
    keep = np.nonzero(np.any(w, axis=0))[0]
    return np.ascontiguousarray(w[:, keep]), np.ascontiguousarray(ind[:, keep])


def imresize(I, scale=None, method='bicubic', sizes=None, mode="vec"):
    """:1268-1306 on a CUDA tensor [H, W] or [H, W, C], uint8 (rounded after each pass) or float64.  The smaller
    scale factor's dimension is resized first, as upstream."""
    E._require_cuda(I, "image")
    if method not in ('bicubic', 'bilinear'):
        raise ValueError("Unidentified method supplied")
    kernel = _cubic if method == 'bicubic' else _triangle
    if I.dtype not in (torch.uint8, torch.float64):
        raise RuntimeError("imresize works on uint8 or float64 images (the reference computes in float64)")
    if scale is not None:
        sc = [float(scale), float(scale)]
        osz = [int(math.ceil(sc[k] * I.shape[k])) for k in range(2)]
    elif sizes is not None:
        sc = [1.0 * sizes[k] / I.shape[k] for k in range(2)]
        osz = list(sizes)
    else:
        raise ValueError("scalar_scale OR output_shape should be defined!")
    B = I.contiguous()
    flag2d = B.dim() == 2
    if flag2d:
        B = B.unsqueeze(2)
    keep = []                                           # tap tables stay referenced until the stream has used them
    for dim in np.argsort(np.array(sc)):
        dim = int(dim)
        w, ind = contributions(I.shape[dim], osz[dim], sc[dim], kernel, 4.0)
        wd = torch.from_numpy(w).to(B.device)
        idd = torch.from_numpy(ind).to(B.device)
        H, W, C_ = B.shape
        oshape = (osz[0], W, C_) if dim == 0 else (H, osz[1], C_)
        out = torch.empty(oshape, dtype=B.dtype, device=B.device)
        L.check(L.lib().srd_resize_dim(B.data_ptr(), 1 if B.dtype == torch.uint8 else 0, H, W, C_, dim, osz[dim],
                                       wd.data_ptr(), idd.data_ptr(), w.shape[1], out.data_ptr(), E._stream()),
                "srd_resize_dim")
        keep.append((wd, idd, B))      # (the caching allocator is stream-ordered: freeing them later is safe, no host sync)
        B = out
    return B.squeeze(2) if flag2d else B


def ssim_uniform_u8(a, b, win_size=7):
    """skimage.metrics.structural_similarity(a, b, multichannel=True, data_range=255) on uint8 [H, W, C] CUDA tensors."""
    E._require_cuda(a, "a")
    if a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.shape != b.shape or a.dim() != 3:
        raise RuntimeError("ssim_uniform_u8 expects two uint8 [H, W, C] tensors of the same shape")
    a, b = a.contiguous(), b.contiguous()
    ws = torch.empty(512, dtype=torch.float64, device=a.device)
    out = torch.empty(1, dtype=torch.float64, device=a.device)
    L.check(L.lib().srd_ssim_uniform_u8(a.data_ptr(), b.data_ptr(), a.shape[0], a.shape[1], a.shape[2], win_size,
                                        ws.data_ptr(), out.data_ptr(), E._stream()), "srd_ssim_uniform_u8")
    return float(out.item())


class Measure:
    """:1499-1542.  `lpips_fn(imgA_hwc_u8, imgB_hwc_u8) -> float` is optional (no pretrained LPIPS weights offline)."""

    def __init__(self, net='alex', lpips_fn=None):
        self.net, self.lpips_fn = net, lpips_fn

    def measure(self, imgA, imgB, img_lr, sr_scale):
        """imgA / imgB / img_lr: [C, H, W] CUDA tensors, uint8 or float in [-1, 1]."""
        if imgA.dtype != torch.uint8:
            imgA, imgB, img_lr = to_uint8_round(imgA), to_uint8_round(imgB), to_uint8_round(img_lr)
        imgA = imgA.permute(1, 2, 0).contiguous()
        imgA_lr = imresize(imgA, 1 / sr_scale)
        imgB = imgB.permute(1, 2, 0).contiguous()
        img_lr = img_lr.permute(1, 2, 0).contiguous()
        res = {'psnr': self.psnr(imgA, imgB), 'ssim': self.ssim(imgA, imgB), 'lr_psnr': self.psnr(imgA_lr, img_lr)}
        if self.lpips_fn is not None:
            res['lpips'] = self.lpips(imgA, imgB)
        return {k: float(v) for k, v in res.items()}

    def lpips(self, imgA, imgB, model=None):
        if self.lpips_fn is None:
            raise NotImplementedError("LPIPS needs pretrained weights that are not available offline; pass lpips_fn")
        return self.lpips_fn(imgA, imgB)

    def ssim(self, imgA, imgB):
        return ssim_uniform_u8(imgA, imgB)

    def psnr(self, imgA, imgB):
        if imgA.dtype != torch.uint8:
            imgA, imgB = to_uint8_round(imgA), to_uint8_round(imgB)
        return psnr_u8(imgA, imgB)
