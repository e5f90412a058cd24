"""B200-native SRDiff reverse-diffusion inference (drop-in for the reference's srdiff/model.py API).

    import dgm_img_super_resolution_b200.model as model
    model.hparams.update(yaml.safe_load(open("srdiff/config.yaml")))   # + unet_dim_mults -> tuple
    rrdb = model.RRDBNet(3, 3, 32, 8, 16); unet = model.Unet(64, out_dim=3, cond_dim=32, dim_mults=(1, 2, 3, 4))
    diffusion = model.GaussianDiffusion(unet, rrdb, timesteps=100).cuda()
    img, rrdb_out = diffusion.sample(img_lr, img_lr_up, img_lr_up.shape)
"""
from . import model  # noqa: F401
from .model import (GaussianDiffusion, RRDBNet, Unet, hparams, runtime)  # noqa: F401

__all__ = ["model", "GaussianDiffusion", "RRDBNet", "Unet", "hparams", "runtime"]
