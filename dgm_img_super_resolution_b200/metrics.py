"""the reference's quality metrics on the GPU (srdiff/model.py:146-222 `SSIM` / `ssim`, and the
PSNR of `Measure` :1499-1542), behind the same names and call signatures.  CUDA tensors only; the arithmetic is
`srd_ssim`, `srd_f32_to_u8_round`, `srd_sq_err_u8` of libsrdiff_b200.so (csrc/variant_ops.cu).

Not built: `Measure.lpips` (needs the pretrained LPIPS/AlexNet weights, which neither the reference nor this image
ships), `Measure.ssim` (skimage's uniform-window SSIM, a different algorithm from the torch `SSIM` above) and
`imresize` (:1268-1306, used for `lr_psnr`) -- calling them raises NotImplementedError.
"""
import math

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


class Measure:
    """:1499-1542 -- the PSNR entry; see the module docstring for what is not built."""

    def __init__(self, net='alex'):
        self.net = net

    def psnr(self, imgA, imgB):
        if imgA.dtype != torch.uint8:
            imgA, imgB = to_uint8_round(imgA), to_uint8_round(imgB)
        return psnr_u8(imgA, imgB)

    def measure(self, imgA, imgB, img_lr, sr_scale):
        raise NotImplementedError("Measure.measure needs LPIPS weights, skimage's SSIM and matlab imresize "
                                  "(SURVEY 8f-4); use Measure.psnr / metrics.SSIM")

    def lpips(self, imgA, imgB, model=None):
        raise NotImplementedError("LPIPS needs pretrained weights that are not available offline")

    def ssim(self, imgA, imgB):
        raise NotImplementedError("skimage's uniform-window SSIM is not built; metrics.SSIM is the torch SSIM :146-222")
