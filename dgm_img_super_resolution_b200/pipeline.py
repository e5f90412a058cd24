"""Image handling around the sampler, on the GPU: the reference driver's bicubic
x`sr_scale` upsample, ToTensor/Normalize, display conversion (st_app1.py:1585-1597, 1661-1686,
1710-1712) and ImgMerger's tile merge (srdiff/model.py:121-145).  Thin Python over the C ABI
(include/srdiff_b200.h); every function requires CUDA tensors -- there is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib as L
from . import engine as E


def _f32(x, what):
    E._require_cuda(x, what)
    return x.contiguous().float()


def _on_device(fn):
    """Run `fn` with its first tensor argument's device current (streams and launches follow the tensor)."""
    import functools

    @functools.wraps(fn)
    def wrapped(x, *a, **k):
        if isinstance(x, torch.Tensor) and x.is_cuda:
            with torch.cuda.device(x.device):
                return fn(x, *a, **k)
        return fn(x, *a, **k)
    return wrapped


@_on_device
def bicubic_upsample(x, scale_factor):
    """F.interpolate(x, scale_factor, mode='bicubic', align_corners=True) (`upsample_tensor`, st_app1.py:1593)."""
    x = _f32(x, "bicubic_upsample input")
    B, C_, H, W = x.shape
    Ho, Wo = int(H * scale_factor), int(W * scale_factor)
    out = torch.empty((B, C_, Ho, Wo), dtype=torch.float32, device=x.device)
    L.check(L.lib().srd_bicubic_upsample(x.data_ptr(), B * C_, H, W, Ho, Wo, out.data_ptr(), E._stream()),
            "srd_bicubic_upsample")
    return out


@_on_device
def normalize_u8(img_u8, mean=None, std=None):
    """uint8 [B, H, W, C] -> float [B, C, H, W]: ToTensor + Normalize(mean, std); without mean/std u/127.5 - 1."""
    E._require_cuda(img_u8, "normalize_u8 input")
    if img_u8.dtype != torch.uint8 or img_u8.dim() != 4:
        raise TypeError("normalize_u8 expects a uint8 [B, H, W, C] tensor")
    img_u8 = img_u8.contiguous()
    B, H, W, C_ = img_u8.shape
    out = torch.empty((B, C_, H, W), dtype=torch.float32, device=img_u8.device)
    m = s = None
    if mean is not None:
        m = torch.tensor(list(mean), dtype=torch.float32, device=img_u8.device)
        s = torch.tensor(list(std), dtype=torch.float32, device=img_u8.device)
    L.check(L.lib().srd_u8hwc_to_nchw(img_u8.data_ptr(), B, H, W, C_, m.data_ptr() if m is not None else None,
                                      s.data_ptr() if s is not None else None, out.data_ptr(), E._stream()),
            "srd_u8hwc_to_nchw")
    return out


@_on_device
def to_uint8(img):
    """float [B, C, H, W] in [-1, 1] -> uint8 [B, H, W, C]: clamp, (x+1)/2, clip(x*255, 0, 255), truncate."""
    img = _f32(img, "to_uint8 input")
    B, C_, H, W = img.shape
    out = torch.empty((B, H, W, C_), dtype=torch.uint8, device=img.device)
    L.check(L.lib().srd_nchw_to_u8hwc(img.data_ptr(), B, C_, H, W, out.data_ptr(), E._stream()), "srd_nchw_to_u8hwc")
    return out


@_on_device
def split_tiles(img, th, tw):
    """[n, C, nx*th, ny*tw] -> ([n*nx*ny, C, th, tw], (nx, ny)); tile (x, y) of image i is entry i*nx*ny + x*ny + y."""
    img = _f32(img, "split_tiles input")
    n, C_, H, W = img.shape
    if H % th or W % tw:
        raise ValueError("image %dx%d is not a whole number of %dx%d tiles" % (H, W, th, tw))
    nx, ny = H // th, W // tw
    out = torch.empty((n * nx * ny, C_, th, tw), dtype=torch.float32, device=img.device)
    L.check(L.lib().srd_tiles(img.data_ptr(), out.data_ptr(), n, C_, nx, ny, th, tw, 1, E._stream()), "srd_tiles")
    return out, (nx, ny)


@_on_device
def merge_tiles(tiles, nx, ny):
    """Inverse of split_tiles; the grid concatenation of ImgMerger.compute (:133-145)."""
    tiles = _f32(tiles, "merge_tiles input")
    N, C_, th, tw = tiles.shape
    if N % (nx * ny):
        raise ValueError("%d tiles do not fill %dx%d grids" % (N, nx, ny))
    n = N // (nx * ny)
    out = torch.empty((n, C_, nx * th, ny * tw), dtype=torch.float32, device=tiles.device)
    L.check(L.lib().srd_tiles(tiles.data_ptr(), out.data_ptr(), n, C_, nx, ny, th, tw, 0, E._stream()), "srd_tiles")
    return out


class ImgMerger:
    """srdiff/model.py:121-145 with the same push()/compute() protocol; `imgs` are lists of [C, h, w] numpy
    arrays or tensors.  The grid merge runs on the GPU (`device`), results go to `eval_fn` as numpy arrays."""

    def __init__(self, eval_fn, device="cuda:0"):
        self.eval_fn, self.device = eval_fn, device
        self.clear()

    def clear(self):
        self.loc2imgs, self.max_x, self.max_y = {}, 0, 0

    def push(self, imgs, loc, loc_bdr):
        self.max_x, self.max_y = loc_bdr
        x, y = loc
        self.loc2imgs['%d,%d' % (x, y)] = imgs
        if len(self.loc2imgs) == self.max_x * self.max_y:
            return self.compute()

    def compute(self):
        img_inputs = []
        for i in range(len(self.loc2imgs['0,0'])):
            first = np.asarray(self.loc2imgs['0,0'][i])
            tiles = torch.stack([torch.as_tensor(np.asarray(self.loc2imgs['%d,%d' % (x, y)][i]), dtype=torch.float32)
                                 for x in range(self.max_x) for y in range(self.max_y)]).to(self.device)
            merged = merge_tiles(tiles, self.max_x, self.max_y)[0].cpu().numpy()
            # np.concatenate keeps the tiles' dtype (uint8 [0, 255] arrays in the reference's use); the fp32 round
            # trip through the merge kernel is exact for every integer / float value a uint8 or fp32 tile holds
            img_inputs.append(merged.astype(first.dtype) if first.dtype != np.float64 else merged.astype(np.float64))
        self.clear()
        return self.eval_fn(*img_inputs)


@torch.no_grad()
def enhance(diffusion, img_lr, sr_scale=None, tile=None):
    """`model_inference` of st_app1.py:1661-1686 for a batch of preprocessed LR images [B, 3, h, w] on the GPU:
    bicubic upsample -> sample() -> uint8 [B, H, W, 3].  tile=(th, tw): the LR image is cut into
    non-overlapping tiles that are sampled as one batch and merged ImgMerger-style (BASELINE configs[2])."""
    from . import model as M
    s = int(sr_scale or M.hparams['sr_scale'])
    img_lr = _f32(img_lr, "enhance input")
    grid = None
    if tile is not None:
        img_lr, grid = split_tiles(img_lr, tile[0], tile[1])
    img_lr_up = bicubic_upsample(img_lr, s)
    out = diffusion.sample(img_lr, img_lr_up, tuple(img_lr_up.shape))[0]
    if grid is not None:
        out = merge_tiles(out, grid[0], grid[1])
    return to_uint8(out)
