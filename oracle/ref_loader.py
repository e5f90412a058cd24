"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference implementation.

Imports ``/root/reference/srdiff/model.py`` in-process so the oracle restatement
(`oracle/srdiff_oracle.py`) can be pinned against the reference's own code and so
golden vectors can be generated (`oracle/make_golden.py`).  ``/root/reference``
only exists in the build container; on the GPU box `available()` is False and
everything falls back to the committed fixtures under ``tests/golden``.

The reference imports four packages that are absent here (matplotlib, pytz,
lpips, skimage -- `srdiff/model.py:23,35,44-46`); none is touched by the hot
path, so they are replaced by empty stub modules before the import.
"""
import importlib.util
import os
import sys
import types
import warnings

import torch
import yaml

REF_ROOT = os.environ.get("SRDIFF_REFERENCE_ROOT", "/root/reference")
_REF = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "srdiff", "model.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def load_reference():
    """Return the reference module object (``srdiff/model.py`` executed as ``ref_model``)."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    _stub("pytz")
    _stub("lpips", LPIPS=object)
    _stub("skimage")
    _stub("skimage.metrics", peak_signal_noise_ratio=None, structural_similarity=None)
    sys.modules["skimage"].metrics = sys.modules["skimage.metrics"]
    spec = importlib.util.spec_from_file_location(
        "ref_model", os.path.join(REF_ROOT, "srdiff", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    _REF = mod
    return mod


def load_hparams(overrides=None) -> dict:
    """`srdiff/config.yaml` the way `st_app1.py:1568-1578` loads it."""
    with open(os.path.join(REF_ROOT, "srdiff", "config.yaml")) as f:
        hp = yaml.safe_load(f)
    if isinstance(hp.get("unet_dim_mults"), str):
        hp["unet_dim_mults"] = tuple(int(v) for v in hp["unet_dim_mults"].split("|"))
    if overrides:
        hp.update(overrides)
    return hp


def build_reference(hp: dict, timesteps=None, seed=1234, nf=None, gc=None, nb=None):
    """Construct RRDBNet/Unet/GaussianDiffusion exactly as upstream's config implies
    (nf=rrdb_num_feat, gc=nf//2, cond_dim=nf)."""
    ref = load_reference()
    ref.hparams.clear()
    ref.hparams.update(hp)
    nf = hp["rrdb_num_feat"] if nf is None else nf
    gc = nf // 2 if gc is None else gc
    nb = hp["rrdb_num_block"] if nb is None else nb
    torch.manual_seed(seed)
    rrdb = ref.RRDBNet(3, 3, nf, nb, gc)
    unet = ref.Unet(hp["hidden_size"], out_dim=3, cond_dim=nf, dim_mults=hp["unet_dim_mults"])
    T = hp["timesteps"] if timesteps is None else timesteps
    diff = ref.GaussianDiffusion(unet, rrdb, timesteps=T, loss_type=hp.get("loss_type", "l1"))
    diff.eval()
    diff.sample_tqdm = False
    return ref, diff


class inject_noise:
    """Context manager: make the reference's sampler consume caller-provided noise.

    RNG draw order in `GaussianDiffusion.sample` (`srdiff/model.py:1070,1057`): one
    ``torch.randn(shape)`` before the RRDB forward, then one ``noise_like`` per step
    (T draws, including t=0).  ``noises`` is a tensor [T+1, *shape] in that order.
    """

    def __init__(self, ref, noises):
        self.ref, self.noises, self.i = ref, noises, 0

    def _next(self):
        n = self.noises[self.i]
        self.i += 1
        return n.clone()

    def __enter__(self):
        self._old_nl = self.ref.noise_like
        self._old_randn = torch.randn
        self.ref.noise_like = lambda shape, device, repeat=False: self._next().to(device)

        def fake_randn(*a, **k):
            return self._next().to(k.get("device", "cpu"))

        self.ref.torch.randn = fake_randn  # same object as torch.randn
        return self

    def __exit__(self, *exc):
        self.ref.noise_like = self._old_nl
        torch.randn = self._old_randn
        return False
