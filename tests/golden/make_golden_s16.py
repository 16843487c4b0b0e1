"""Generates tests/golden/s16_train_color_b8.npz with the fp64 CPU oracle: one adversarial.lua iteration on the
--scale 16 nets (models.lua:27-51, :279-316), composed by tests/s16_utils.oracle_iteration.  Like the other goldens it
pins the oracle restatement (the reference ships no vectors and cannot run here); inputs are regenerated from the seed.
Run:  python tests/golden/make_golden_s16.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import s16_utils as SU  # noqa: E402

STRIDE = 1009


def sample(v):
    return np.asarray(v, np.float64).ravel()[::STRIDE].copy()


def main():
    B, C, seed, init = 8, 3, 4242, "near"
    case = SU.make_case(B, C, seed=seed, init=init)
    res = SU.oracle_iteration(case, B, C)
    out = dict(B=B, C=C, seed=seed, init=init, input_checksum=float(sum(np.abs(case[k]).sum() for k in sorted(case))),
               lossD=res["lossD"], lossG=res["lossG"], conf=res["conf"], gradD=sample(res["gradD"]), gradG=sample(res["gradG"]),
               gradD_absmax=np.abs(res["gradD"]).max(), gradG_absmax=np.abs(res["gradG"]).max(), PD=sample(res["PD"]),
               PG=sample(res["PG"]), fake=res["fake"].astype(np.float32), outD=res["outD"], outG=res["outG"], bn=res["bn"])
    np.savez_compressed(os.path.join(HERE, "s16_train_color_b8.npz"), **out)
    print("s16_train_color_b8 lossD", res["lossD"], "lossG", res["lossG"])


if __name__ == "__main__":
    main()
