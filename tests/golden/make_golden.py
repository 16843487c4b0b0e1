"""Generates the committed golden vectors in tests/golden/*.npz with the fp64 CPU oracle.

The reference ships no golden vectors and cannot run in this image, so these pins are
produced by the oracle restatement (cross-checked against PyTorch-CPU in tests/test_oracle_vs_torch.py).
Inputs are regenerated from seeds by tests/parity_utils.make_case; only outputs / strided samples are
stored so the files stay small.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import parity_utils as PU  # noqa: E402
from oracle import oracle as O  # noqa: E402

STRIDE = 1009  # prime stride for gradient / parameter samples


def sample(v):
    return np.asarray(v, np.float64).ravel()[::STRIDE].copy()


def train_case(name, B, C, seed, init, iters=2):
    case = PU.make_case(B, C, seed=seed, init=init)
    st = PU.fresh_state(case)
    out = {"B": B, "C": C, "seed": seed, "init": init, "input_checksum": float(sum(np.abs(case[k]).sum() for k in sorted(case)))}
    for it in range(iters):
        res = O.f64.train_iteration(B, C, PU.HYPER, case["real"], case["noise_D"], case["noise_G"], case["masks_D"],
                                    case["masks_G"], st)
        p = "it%d_" % it
        out[p + "lossD"], out[p + "lossG"], out[p + "conf"] = res["lossD"], res["lossG"], res["conf"]
        out[p + "gradD"], out[p + "gradG"] = sample(res["gradD"]), sample(res["gradG"])
        out[p + "gradD_absmax"], out[p + "gradG_absmax"] = np.abs(res["gradD"]).max(), np.abs(res["gradG"]).max()
        out[p + "PD"], out[p + "PG"] = sample(st["PD"]), sample(st["PG"])
        out[p + "mD"], out[p + "mG"] = sample(st["mD"]), sample(st["mG"])
        out[p + "fake"], out[p + "outD"] = res["fake"].astype(np.float32), res["outD"]
        out[p + "bnG"] = st["bnG"].copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "lossD", out["it0_lossD"], "lossG", out["it0_lossG"])


def nets_case(name, B, C, seed):
    case = PU.make_case(2 * B, C, seed=seed)
    rng = np.random.default_rng(seed + 1)
    noise = case["noise_G"][:B]
    dout = rng.standard_normal((B, C, 32, 32)).astype(np.float32)
    g = O.f64.G()
    img = g.forward(case["PG"], noise, C)
    dP, dn = g.backward(dout, want_dnoise=True)
    x = rng.random((B, C, 32, 32)).astype(np.float32)
    dd = rng.standard_normal(B).astype(np.float32)
    d = O.f64.D()
    out_d = d.forward(case["PD"], x, case["masks_D"][:B])
    dPD, dimg = d.backward(dd)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), B=B, C=C, seed=seed, G_out=img.astype(np.float32),
                        G_dP=sample(dP), G_dP_absmax=np.abs(dP).max(), G_dnoise=dn, D_out=out_d, D_dP=sample(dPD),
                        D_dP_absmax=np.abs(dPD).max(), D_dimg=dimg.astype(np.float32),
                        z1=sample(g.tap("z1")), z2=sample(g.tap("z2")))
    print(name, "G_out mean", img.mean(), "D_out", out_d[:3])


def ops_case(name):
    rng = np.random.default_rng(77)
    x = rng.standard_normal((2, 6, 8, 8))
    w = rng.standard_normal((5, 6, 5, 5)) * 0.2
    b = rng.standard_normal(5)
    dy = rng.standard_normal((2, 5, 8, 8))
    out = {"conv_y": O.f64.conv_fwd(x, w, b)}
    out["conv_dx"], out["conv_dw"], out["conv_db"] = O.f64.conv_bwd(x, w, dy)
    g, be = rng.uniform(0.5, 1.5, 6), rng.standard_normal(6)
    out["bn_y"], out["bn_mean"], out["bn_istd"] = O.f64.bn_fwd_train(x, g, be)
    out["bn_dx"], out["bn_dg"], out["bn_db"] = O.f64.bn_bwd(x, g, out["bn_mean"], out["bn_istd"], rng.standard_normal(x.shape))
    out["up"] = O.f64.up2_fwd(x)
    out["pool"] = O.f64.avgpool2_fwd(x)
    xs = rng.uniform(0.01, 0.99, 16)
    ts = (rng.random(16) < 0.5).astype(np.float64)
    out["bce"], out["bce_grad"] = O.f64.bce_fwd(xs, ts), O.f64.bce_bwd(xs, ts)
    p, gr, m, v = rng.standard_normal(50), rng.standard_normal(50), np.zeros(50), np.zeros(50)
    for t in (1, 2, 3):
        O.f64.adam(p, gr, m, v, t)
    out["adam_p"], out["adam_m"], out["adam_v"] = p, m, v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


if __name__ == "__main__":
    ops_case("ops_small")
    nets_case("nets_b4_c3", 4, 3, 301)
    nets_case("nets_b4_c1", 4, 1, 302)
    train_case("train_config1_gray_b16", 16, 1, 401, "trained")  # BASELINE.json configs[0]
    train_case("train_color_b8_refinit", 8, 3, 402, "reference", iters=1)
