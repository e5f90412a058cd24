import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SEED = 1234


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_hparams():
    with open(os.path.join(GOLDEN, "hparams.json")) as f:
        hp = json.load(f)
    hp["unet_dim_mults"] = tuple(hp["unet_dim_mults"])
    return hp


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=True)


def build_ours(hp=None, timesteps=None, seed=SEED, device=None):
    """Our modules with the reference's seeded default init (identical weights, see test_api)."""
    import dgm_img_super_resolution_b200.model as M
    hp = hp or load_hparams()
    M.hparams.clear()
    M.hparams.update(hp)
    torch.manual_seed(seed)
    nf = hp["rrdb_num_feat"]
    rrdb = M.RRDBNet(3, 3, nf, hp["rrdb_num_block"], nf // 2)
    unet = M.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
    diff = M.GaussianDiffusion(unet, rrdb, timesteps=hp["timesteps"] if timesteps is None else timesteps,
                               loss_type=hp["loss_type"])
    diff.eval()
    diff.sample_tqdm = False
    if device is not None:
        diff = diff.to(device)
    return M, diff


def gen_inputs(B, h, w, scale, seed):
    """Same generator protocol as oracle/make_golden.py."""
    g = torch.Generator().manual_seed(seed)
    img_lr = torch.rand((B, 3, h, w), generator=g) * 2 - 1
    img_lr_up = torch.nn.functional.interpolate(img_lr, scale_factor=scale, mode="bicubic", align_corners=True)
    return img_lr, img_lr_up, g


@pytest.fixture(scope="session")
def hp():
    return load_hparams()


@pytest.fixture(scope="session")
def cpu_state(hp):
    """(state_dict, schedule) on the CPU for the oracle, from our seeded containers."""
    M, diff = build_ours(hp)
    sd = {k: v.clone() for k, v in diff.state_dict().items()}
    return sd
