"""Generates tests/golden/c2f_*.npz with the fp64 CPU oracle (fg_oracle_c2f.h).  The reference ships no vectors for
train_c2f.lua either; inputs are regenerated from seeds by tests/c2f_utils.make_case, only outputs / strided samples
are stored.  Run:  python tests/golden/make_golden_c2f.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import c2f_utils as CU  # noqa: E402

STRIDE = 1009


def sample(v):
    return np.asarray(v, np.float64).ravel()[::STRIDE].copy()


def train_case(name, B, C, seed, init):
    case = CU.make_case(B, C, seed=seed, init=init)
    res = CU.oracle_iteration(case, B, C)
    st = res["state"]
    out = dict(B=B, C=C, seed=seed, init=init, input_checksum=float(sum(np.abs(case[k]).sum() for k in sorted(case))),
               lossD=res["lossD"], lossG=res["lossG"], conf=res["conf"], gradD=sample(res["gradD"]),
               gradG=sample(res["gradG"]), gradD_absmax=np.abs(res["gradD"]).max(), gradG_absmax=np.abs(res["gradG"]).max(),
               PD=sample(st["PD"]), PG=sample(st["PG"]), mD=sample(st["mD"]), mG=sample(st["mG"]),
               fake=res["fake"].astype(np.float32), outD=res["outD"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "lossD", out["lossD"], "lossG", out["lossG"], "conf", out["conf"])


if __name__ == "__main__":
    train_case("c2f_train_color_b8", 8, 3, 501, "trained")
    train_case("c2f_train_gray_b4_smooth", 4, 1, 502, "smooth")
