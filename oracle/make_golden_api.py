"""TEST INFRASTRUCTURE ONLY -- fixtures for the stand-alone module forwards and sampler helper methods of the
reference (`/root/reference/srdiff/model.py:398-434, 670-706, 962-990`), produced by running the UNMODIFIED
reference in the build container.

    python -m oracle.make_golden_api      ->  tests/golden/api_rows.pt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import OUT, SEED  # noqa: E402


def main():
    hp = ref_loader.load_hparams()
    ref, diff = ref_loader.build_reference(hp, seed=SEED)
    g = torch.Generator().manual_seed(29)
    out = {}
    with torch.no_grad():
        # sampler helpers (per-sample t)
        x = torch.randn((4, 3, 8, 12), generator=g)
        e = torch.randn((4, 3, 8, 12), generator=g)
        t = torch.tensor([0, 1, 50, 99], dtype=torch.long)
        mean, var, logvar = diff.q_mean_variance(x, t)
        x0 = diff.predict_start_from_noise(x, t, e)
        pm, pv, plv = diff.q_posterior(e, x, t)
        mm, mv, mlv, xr = diff.p_mean_variance(x.clone(), t, e, True)
        mm2, _, _, xr2 = diff.p_mean_variance(x.clone(), t, e, False)
        out["helpers"] = {"x": x, "eps": e, "t": t, "q_mean": mean, "q_var": var, "q_logvar": logvar,
                          "x0": x0, "post_mean": pm, "post_var": pv, "post_logvar": plv,
                          "pmv_mean": mm, "pmv_var": mv, "pmv_logvar": mlv, "pmv_x0": xr,
                          "pmv_mean_noclip": mm2, "pmv_x0_noclip": xr2}
        # SinusoidalPosEmb
        tt = torch.tensor([0., 1., 17., 99., 999.])
        out["pos_emb"] = {"t": tt, "dim": 64, "out": ref.SinusoidalPosEmb(64)(tt)}
        # dense blocks of the encoder (first trunk block of the seeded RRDBNet)
        blk = diff.rrdb.RRDB_trunk[0]
        xb = torch.randn((2, hp["rrdb_num_feat"], 10, 6), generator=g)
        out["rdb"] = {"x": xb, "rdb1": blk.RDB1(xb), "rrdb": blk(xb)}
        # Residual / Rezero around a parameter-free function
        xm = torch.randn((2, 5, 4, 4), generator=g)
        rz = ref.Rezero(ref.Mish())
        rz.g.data.fill_(0.37)
        out["wrappers"] = {"x": xm, "residual_mish": ref.Residual(ref.Mish())(xm), "rezero_mish_g037": rz(xm)}
    p = os.path.join(OUT, "api_rows.pt")
    torch.save(out, p)
    print(p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
