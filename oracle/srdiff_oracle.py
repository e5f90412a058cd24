"""TEST INFRASTRUCTURE ONLY -- CPU oracle (fp32) for SRDiff reverse-diffusion inference.

A functional restatement of the reference's hot path, written against a flat
``state_dict`` (the key contract of SURVEY Appendix C) instead of nn.Modules.
Every function cites the reference lines (``/root/reference/srdiff/model.py``) it
follows.  Arithmetic is delegated to PyTorch's CPU ATen ops -- exactly the
third-party dependency the reference itself uses (unpinned there; torch 2.11.0
here) -- so, on the same inputs, this oracle reproduces the reference to fp32
round-off (pinned by ``tests/test_oracle.py`` against the live reference when
``/root/reference`` is present, and against the committed fixtures in
``tests/golden`` generated from it by ``oracle/make_golden.py``).

Parity status: the reference ships no tests / golden vectors of its own for this
path, so the pin is "outputs of the reference itself run here".
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- basics
def mish(x):
    """`Mish.forward` :422-424 -- x * tanh(softplus(x)) (torch softplus threshold 20)."""
    return x * torch.tanh(F.softplus(x))


def _weight(sd, key):
    """Effective conv weight: plain, or under `torch.nn.utils.weight_norm` (`apply_weight_norm` :808-814, state
    dict keys weight_g / weight_v) g * v / ||v|| with the norm over every dim but 0."""
    if (key + ".weight") in sd:
        return sd[key + ".weight"]
    v, g = sd[key + ".weight_v"], sd[key + ".weight_g"]
    return v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1, 1))


def _conv(sd, key, x, stride=1, padding=0):
    return F.conv2d(x, _weight(sd, key), sd[key + ".bias"], stride=stride, padding=padding)


def _rpad(x):
    """`nn.ReflectionPad2d(1)` :444,450,498."""
    return F.pad(x, (1, 1, 1, 1), mode="reflect")


# ----------------------------------------------------------------------------- schedule (row G)
def cosine_betas(timesteps, s=0.008):
    """`cosine_beta_schedule` :898-908 (note linspace(0, steps, steps))."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


def _warmup(beta_start, beta_end, T, frac):
    betas = beta_end * np.ones(T, dtype=np.float64)
    w = int(T * frac)
    betas[:w] = np.linspace(beta_start, beta_end, w, dtype=np.float64)
    return betas


def named_betas(T, kind="linear", beta_start=0.0001, beta_end=0.02):
    """`get_beta_schedule` :879-895."""
    if kind == "quad":
        return np.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=np.float64) ** 2
    if kind == "linear":
        return np.linspace(beta_start, beta_end, T, dtype=np.float64)
    if kind == "warmup10":
        return _warmup(beta_start, beta_end, T, 0.1)
    if kind == "warmup50":
        return _warmup(beta_start, beta_end, T, 0.5)
    if kind == "const":
        return beta_end * np.ones(T, dtype=np.float64)
    if kind == "jsd":
        return 1.0 / np.linspace(T, 1, T, dtype=np.float64)
    raise NotImplementedError(kind)


def make_schedule(timesteps, hp):
    """The 12 fp32 buffers of `GaussianDiffusion.__init__` :921-959 (float64 math -> fp32)."""
    if hp["beta_schedule"] == "cosine":
        betas = cosine_betas(timesteps, s=hp["beta_s"])
    elif hp["beta_schedule"] == "linear":
        betas = named_betas(timesteps, beta_end=hp["beta_end"])
        if hp["res"]:
            betas[-1] = 0.999
    else:
        raise NotImplementedError(hp["beta_schedule"])
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {
        "betas": f32(betas),
        "alphas_cumprod": f32(ac),
        "alphas_cumprod_prev": f32(ac_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "log_one_minus_alphas_cumprod": f32(np.log(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
        "posterior_variance": f32(pv),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(pv, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    }


# ----------------------------------------------------------------------------- RRDB encoder (row R)
def _rdb5c(sd, p, x):
    """`ResidualDenseBlock_5C.forward` :684-690."""
    lr = lambda v: F.leaky_relu(v, 0.2)
    x1 = lr(_conv(sd, p + "conv1", x, padding=1))
    x2 = lr(_conv(sd, p + "conv2", torch.cat((x, x1), 1), padding=1))
    x3 = lr(_conv(sd, p + "conv3", torch.cat((x, x1, x2), 1), padding=1))
    x4 = lr(_conv(sd, p + "conv4", torch.cat((x, x1, x2, x3), 1), padding=1))
    x5 = _conv(sd, p + "conv5", torch.cat((x, x1, x2, x3, x4), 1), padding=1)
    return x5 * 0.2 + x


def rrdb_forward(sd, x, hp, prefix="rrdb."):
    """`RRDBNet.forward(x, get_fea=True)` :727-749 -> (out, feas[nb+1])."""
    nb = 0
    while (prefix + "RRDB_trunk.%d.RDB1.conv1.weight" % nb) in sd:
        nb += 1
    lr = lambda v: F.leaky_relu(v, 0.2)
    feas = []
    x = (x + 1) / 2
    fea_first = fea = _conv(sd, prefix + "conv_first", x, padding=1)
    for i in range(nb):
        p = prefix + "RRDB_trunk.%d." % i
        out = _rdb5c(sd, p + "RDB1.", fea)
        out = _rdb5c(sd, p + "RDB2.", out)
        out = _rdb5c(sd, p + "RDB3.", out)
        fea = out * 0.2 + fea                       # `RRDB.forward` :702-706
        feas.append(fea)
    trunk = _conv(sd, prefix + "trunk_conv", fea, padding=1)
    fea = fea_first + trunk
    feas.append(fea)
    up = lambda v: F.interpolate(v, scale_factor=2, mode="nearest")
    fea = lr(_conv(sd, prefix + "upconv1", up(fea), padding=1))
    fea = lr(_conv(sd, prefix + "upconv2", up(fea), padding=1))
    if hp["sr_scale"] == 8:
        fea = lr(_conv(sd, prefix + "upconv3", up(fea), padding=1))
    fea_hr = _conv(sd, prefix + "HRconv", fea, padding=1)
    out = _conv(sd, prefix + "conv_last", lr(fea_hr), padding=1)
    out = out.clamp(0, 1) * 2 - 1
    return out, feas


# ----------------------------------------------------------------------------- Unet denoiser (row U)
def time_embedding(sd, time, dim, prefix="denoise_fn."):
    """`SinusoidalPosEmb` :412-419 followed by `Unet.mlp` :764-768."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, device=time.device) * -e)     # on the input's device, like :415-416
    e = time[:, None] * e[None, :]
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    h = F.linear(e, sd[prefix + "mlp.0.weight"], sd[prefix + "mlp.0.bias"])
    return F.linear(mish(h), sd[prefix + "mlp.2.weight"], sd[prefix + "mlp.2.bias"])


def _block(sd, p, x, groups=0):
    """`Block` :439-457: reflect-pad, conv3x3, [GroupNorm(groups) when groups > 0 :448-454], Mish."""
    h = _conv(sd, p + "block.1", _rpad(x))
    if groups:
        h = F.group_norm(h, groups, sd[p + "block.2.weight"], sd[p + "block.2.bias"], eps=1e-5)
    return mish(h)


def resnet_block(sd, p, x, time_emb=None, cond=None, groups=0):
    """Stand-alone `ResnetBlock.forward(x, time_emb, cond)` :473-480 (the caller's `time_emb` goes through
    `self.mlp` = Mish + Linear :464-468)."""
    h = _block(sd, p + "block1.", x, groups)
    if time_emb is not None:
        h = h + F.linear(mish(time_emb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])[:, :, None, None]
    if cond is not None:
        h = h + cond
    h = _block(sd, p + "block2.", h, groups)
    if (p + "res_conv.bias") in sd:
        return h + _conv(sd, p + "res_conv", x)
    return h + x


def _resnet(sd, p, x, temb):
    """`ResnetBlock.forward` :473-480 (temb added AFTER block1's Mish)."""
    h = _block(sd, p + "block1.", x)
    h = h + F.linear(mish(temb), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"])[:, :, None, None]
    h = _block(sd, p + "block2.", h)
    if (p + "res_conv.bias") in sd:
        return h + _conv(sd, p + "res_conv", x)
    return h + x


def linear_attention(sd, p, x, heads=4):
    """`LinearAttention.forward` :514-522 with einops' rearrange spelled out:
    'b (qkv heads c) h w -> qkv b heads c (h w)' and back 'b heads c (h w) -> b (heads c) h w'."""
    b, _, h, w = x.shape
    qkv = F.conv2d(x, sd[p + "to_qkv.weight"])
    q, k, v = qkv.reshape(b, 3, heads, -1, h * w).unbind(1)
    k = k.softmax(dim=-1)
    context = torch.einsum('bhdn,bhen->bhde', k, v)
    out = torch.einsum('bhde,bhdn->bhen', context, q).reshape(b, -1, h, w)
    return _conv(sd, p + "to_out", out)


def unet_forward(sd, x, time, cond_feas, hp, prefix="denoise_fn.", taps=None, img_lr_up=None):
    """`Unet.forward` :820-847.  `cond_feas` is the RRDB feature list; `taps`, if a dict,
    receives intermediate activations (test instrumentation only)."""
    dim = sd[prefix + "mlp.2.weight"].shape[0]
    n_levels = 0
    while (prefix + "downs.%d.0.block1.block.1.bias" % n_levels) in sd:
        n_levels += 1
    temb = time_embedding(sd, time.float(), dim, prefix)
    s = hp["sr_scale"]
    cond = F.conv_transpose2d(torch.cat(cond_feas[2::3], 1), sd[prefix + "cond_proj.weight"],
                              sd[prefix + "cond_proj.bias"], stride=s, padding=s // 2)
    hs = []
    for i in range(n_levels):
        p = prefix + "downs.%d." % i
        x = _resnet(sd, p + "0.", x, temb)
        x = _resnet(sd, p + "1.", x, temb)
        if i == 0:
            x = x + cond
            if hp["res"] and hp["up_input"]:             # :831-832
                x = x + _conv(sd, prefix + "up_proj.1", _rpad(img_lr_up))
        hs.append(x)
        if i < n_levels - 1:                         # `Downsample` :494-503
            x = _conv(sd, p + "2.conv.1", _rpad(x), stride=2)
        if taps is not None:
            taps["down%d" % i] = x
    x = _resnet(sd, prefix + "mid_block1.", x, temb)
    if hp["use_attn"]:                               # Residual(Rezero(LinearAttention)) :788-789, 836-837
        x = linear_attention(sd, prefix + "mid_attn.fn.fn.", x) * sd[prefix + "mid_attn.fn.g"] + x
    x = _resnet(sd, prefix + "mid_block2.", x, temb)
    if taps is not None:
        taps["mid"] = x
    for j in range(n_levels - 1):                    # three Upsamples, h[0] never popped (Q5)
        p = prefix + "ups.%d." % j
        x = torch.cat((x, hs.pop()), dim=1)
        x = _resnet(sd, p + "0.", x, temb)
        x = _resnet(sd, p + "1.", x, temb)
        x = F.conv_transpose2d(x, sd[p + "2.conv.0.weight"], sd[p + "2.conv.0.bias"],
                               stride=2, padding=1)  # `Upsample` :483-491
        if taps is not None:
            taps["up%d" % j] = x
    x = _block(sd, prefix + "final_conv.0.", x)
    return _conv(sd, prefix + "final_conv.1", x)


# ----------------------------------------------------------------------------- sampler (rows P, S, F)
def _ext(a, t, ndim):
    """`extract` :860-863."""
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))


def p_sample_update(sched, x, t, eps, noise, clip_denoised=True):
    """The arithmetic of `p_sample` :1054-1060 given eps = Unet(x, t):
    `predict_start_from_noise` :968-972, clamp :986-987, `q_posterior` :974-981."""
    nd = x.dim()
    x0 = _ext(sched["sqrt_recip_alphas_cumprod"], t, nd) * x - \
        _ext(sched["sqrt_recipm1_alphas_cumprod"], t, nd) * eps
    if clip_denoised:
        x0 = x0.clamp(-1.0, 1.0)
    mean = _ext(sched["posterior_mean_coef1"], t, nd) * x0 + \
        _ext(sched["posterior_mean_coef2"], t, nd) * x
    logvar = _ext(sched["posterior_log_variance_clipped"], t, nd)
    mask = (1 - (t == 0).float()).reshape(x.shape[0], *((1,) * (nd - 1)))
    return mean + mask * (0.5 * logvar).exp() * noise, x0


def res2img(img, img_lr_up, hp, clip_input=None):
    """`res2img` :1118-1125."""
    if clip_input is None:
        clip_input = hp["clip_input"]
    if hp["res"]:
        if clip_input:
            img = img.clamp(-1, 1)
        img = img / hp["res_rescale"] + img_lr_up
    return img


def sample(sd, sched, hp, img_lr, img_lr_up, noises, record=None):
    """`GaussianDiffusion.sample` :1062-1091 with caller-provided noise.

    ``noises`` is [T+1, B, 3, H, W]: noises[0] is x_T (the `torch.randn(shape)` at :1070),
    noises[1 + k] the k-th per-step draw (step t = T-1-k; the t == 0 draw is multiplied by 0).
    ``record`` (a dict) receives per-step eps and x lists.  Returns (img, rrdb_out).
    """
    T = sched["betas"].shape[0]
    b = img_lr.shape[0]
    x = noises[0].clone()
    if hp["use_rrdb"]:
        rrdb_out, cond = rrdb_forward(sd, img_lr, hp)
    else:
        raise NotImplementedError("use_rrdb=False is outside the hot path")
    for k, i in enumerate(reversed(range(T))):
        t = torch.full((b,), i, dtype=torch.long)
        eps = unet_forward(sd, x, t, cond, hp, img_lr_up=img_lr_up)
        x, x0 = p_sample_update(sched, x, t, eps, noises[1 + k])
        if record is not None:
            record.setdefault("eps", []).append(eps)
            record.setdefault("x", []).append(x)
            record.setdefault("x0", []).append(x0)
    return res2img(x, img_lr_up, hp), rrdb_out



# ----------------------------------------------------------------------------- SURVEY 8f "next" rows
def img2res(x, img_lr_up, hp, clip_input=None):
    """`img2res` :1127-1134."""
    if clip_input is None:
        clip_input = hp["clip_input"]
    if hp["res"]:
        x = (x - img_lr_up) * hp["res_rescale"]
        if clip_input:
            x = x.clamp(-1, 1)
    return x


def q_sample(sched, x_start, t, noise):
    """`q_sample` :1041-1048 (t < 0 returns x_start)."""
    nd = x_start.dim()
    t_cond = (t[:, None, None, None] >= 0).float()
    tc = t.clamp_min(0)
    return (_ext(sched["sqrt_alphas_cumprod"], tc, nd) * x_start +
            _ext(sched["sqrt_one_minus_alphas_cumprod"], tc, nd) * noise) * t_cond + x_start * (1 - t_cond)


def interpolate(sd, sched, hp, x1, x2, img_lr, img_lr_up, noises, t, lam):
    """`interpolate` :1093-1116; noises[0:2] feed the two q_sample calls, noises[2 + k] the k-th p_sample."""
    b = x1.shape[0]
    _, cond = rrdb_forward(sd, img_lr, hp)
    r1, r2 = img2res(x1, img_lr_up, hp), img2res(x2, img_lr_up, hp)
    tb = torch.full((b,), t, dtype=torch.long)
    xt1, xt2 = q_sample(sched, r1, tb, noises[0]), q_sample(sched, r2, tb, noises[1])
    x = (1 - lam) * xt1 + lam * xt2
    for k, i in enumerate(reversed(range(0, t))):
        ti = torch.full((b,), i, dtype=torch.long)
        x, _ = p_sample_update(sched, x, ti, unet_forward(sd, x, ti, cond, hp), noises[2 + k])
    return res2img(x, img_lr_up, hp)


def sample_nores(sd, sched, hp, img_lr, img_lr_up, noises):
    """`sample` :1062-1091 with hparams['res'] = False: start = q_sample(img_lr_up, T-1) :1066-1068."""
    T = sched["betas"].shape[0]
    b = img_lr.shape[0]
    x = q_sample(sched, img_lr_up, torch.full((b,), T - 1, dtype=torch.long), noises[0])
    rrdb_out, cond = rrdb_forward(sd, img_lr, hp)
    for k, i in enumerate(reversed(range(T))):
        t = torch.full((b,), i, dtype=torch.long)
        x, _ = p_sample_update(sched, x, t, unet_forward(sd, x, t, cond, hp), noises[1 + k])
    return res2img(x, img_lr_up, hp), rrdb_out


def bicubic_up(x, scale):
    """`upsample_tensor` st_app1.py:1593-1597 (the same third-party ATen op the reference calls)."""
    return F.interpolate(x, scale_factor=scale, mode="bicubic", align_corners=True)


def merge_tiles(tiles, nx, ny):
    """`ImgMerger.compute` :133-145 for one image list: tiles[x*ny + y] is [C, th, tw]."""
    rows = [np.concatenate([tiles[x * ny + y] for y in range(ny)], 2) for x in range(nx)]
    return np.concatenate(rows, 1)


def to_uint8_hwc(img):
    """st_app1.py:1682-1684, 1710-1711 for a batch [B, C, H, W] in [-1, 1] -> uint8 [B, H, W, C]."""
    v = (img.clamp(-1, 1) + 1) / 2
    return torch.from_numpy(np.clip(v.permute(0, 2, 3, 1).numpy() * 255, 0, 255).astype(np.uint8))


def normalize_u8(u8_hwc, mean=None, std=None):
    """ToTensor + Normalize, st_app1.py:1585-1589 ([B, H, W, C] uint8 -> [B, C, H, W] float); without
    mean/std the [-1, 1] convention u / 127.5 - 1 of srdiff/model.py:95."""
    x = u8_hwc.permute(0, 3, 1, 2).float()
    if mean is None:
        return x / 127.5 - 1
    m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
    s_ = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
    return (x / 255.0 - m) / s_

# ----------------------------------------------------------------------------- helpers for tests / bench
def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a, b, peak=2.0):
    """PSNR for images in [-1, 1] (peak-to-peak 2)."""
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)


# ----------------------------------------------------------------------------- metrics
def ssim_torch(img1, img2, window_size=11, size_average=True, map01=True):
    """`SSIM.forward` :196-209 (map01: img * 0.5 + 0.5 first) / functional `ssim` :212-221 + `_ssim` :161-182."""
    if map01:
        img1, img2 = img1 * 0.5 + 0.5, img2 * 0.5 + 0.5
    ch = img1.shape[1]
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    window = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(ch, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=ch)
    mu2 = F.conv2d(img2, window, padding=pad, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, window, padding=pad, groups=ch) - mu1_sq
    s2 = F.conv2d(img2 * img2, window, padding=pad, groups=ch) - mu2_sq
    s12 = F.conv2d(img1 * img2, window, padding=pad, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)


def to_uint8_round(img):
    """`Measure.measure` :1516-1518: np.round((img + 1) * 127.5).clip(0, 255).astype(np.uint8)."""
    return np.round((img.numpy() + 1) * 127.5).clip(min=0, max=255).astype(np.uint8)


def psnr_u8(a, b):
    """`Measure.psnr` :1541-1542 = skimage.metrics.peak_signal_noise_ratio(a, b, data_range=255) (scikit-image, not
    vendored by the reference and absent here; published definition: 10 log10(R^2 / mean((a - b)^2)) in float64)."""
    err = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2, dtype=np.float64)
    return float("inf") if err == 0 else 10 * np.log10((255 ** 2) / err)


def _resize_taps(n_in, n_out, scale, kernel, support=4.0):
    """`contributions` :1198-1218, one output sample at a time: (weights, mirrored indices) lists."""
    widen = scale < 1
    width = support / scale if widen else support
    P = int(math.ceil(width)) + 2
    rows = []
    for o in range(1, n_out + 1):
        u = o / scale + 0.5 * (1 - 1 / scale)
        left = math.floor(u - width / 2)
        idx = np.array([left + p - 1 for p in range(P)], dtype=np.int64)
        d = u - idx - 1
        w = scale * kernel(scale * d) if widen else kernel(d)
        w = w / w.sum()
        m = np.mod(idx, 2 * n_in)
        idx = np.where(m < n_in, m, 2 * n_in - 1 - m)          # mirror at the borders
        rows.append((w, idx))
    used = np.any(np.stack([r[0] for r in rows]) != 0, axis=0)
    return [(w[used], i[used]) for w, i in rows]


def _k_cubic(x):
    """`cubic` :1188-1195."""
    a = np.abs(x)
    return np.where(a <= 1, 1.5 * a ** 3 - 2.5 * a ** 2 + 1, np.where(a <= 2, -0.5 * a ** 3 + 2.5 * a ** 2 - 4 * a + 2, 0.0))


def _k_triangle(x):
    """`triangle` :1179-1185."""
    return np.where((x >= -1) & (x < 0), x + 1, np.where((x >= 0) & (x <= 1), 1 - x, 0.0))


def imresize_np(img, scale=None, sizes=None, method="bicubic"):
    """`imresize` :1268-1306 (`imresizevec` :1240-1252) for [H, W(, C)] uint8 / float64 numpy arrays: separable,
    smaller scale factor first, float64 accumulation, uint8 clipped + rounded half-to-even after each pass."""
    kernel = _k_cubic if method == "bicubic" else _k_triangle
    if scale is not None:
        sc = [float(scale)] * 2
        osz = [int(math.ceil(sc[k] * img.shape[k])) for k in range(2)]
    else:
        sc = [sizes[k] / img.shape[k] for k in range(2)]
        osz = list(sizes)
    B = img if img.ndim == 3 else img[:, :, None]
    for dim in np.argsort(np.array(sc)):
        taps = _resize_taps(img.shape[dim], osz[dim], sc[dim], kernel)
        shape = list(B.shape)
        shape[dim] = osz[dim]
        out = np.zeros(shape, dtype=np.float64)
        for o, (w, idx) in enumerate(taps):
            for wp, ip in zip(w, idx):
                if dim == 0:
                    out[o] += wp * B[ip].astype(np.float64)
                else:
                    out[:, o] += wp * B[:, ip].astype(np.float64)
        B = np.around(np.clip(out, 0, 255)).astype(np.uint8) if img.dtype == np.uint8 else out
    return B if img.ndim == 3 else B[:, :, 0]


def ssim_uniform_u8(a, b, win=7):
    """`Measure.ssim` :1537-1539 = skimage.metrics.structural_similarity(a, b, multichannel=True, data_range=255)
    (scikit-image: not vendored by the reference, absent here -- PARITY UNPINNED for this function; published
    algorithm: Wang et al. 2004 with a uniform win x win window, sample covariance, K1 = 0.01, K2 = 0.03, mean over
    the map cropped by (win - 1) / 2, per channel, then averaged)."""
    from scipy.ndimage import uniform_filter
    pad, NP = (win - 1) // 2, win * win
    cov = NP / (NP - 1.0)
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    vals = []
    for c in range(a.shape[2]):
        X, Y = a[:, :, c].astype(np.float64), b[:, :, c].astype(np.float64)
        ux, uy = uniform_filter(X, size=win), uniform_filter(Y, size=win)
        vx = cov * (uniform_filter(X * X, size=win) - ux * ux)
        vy = cov * (uniform_filter(Y * Y, size=win) - uy * uy)
        vxy = cov * (uniform_filter(X * Y, size=win) - ux * uy)
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        vals.append(S[pad:S.shape[0] - pad, pad:S.shape[1] - pad].mean(dtype=np.float64))
    return float(np.mean(vals))
