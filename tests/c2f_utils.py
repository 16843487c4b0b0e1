"""Seeded cases for the coarse-to-fine parity tests (BASELINE.json configs[3], train_c2f.lua)."""
import numpy as np

from face_generator_b200 import layouts as LY
from oracle import oracle_c2f as OC

# train_c2f.lua:26-34 defaults: D_L1 = 1e-7 is active, everything else off; clamps 1 / 5
HYPER = dict(lr_D=1e-3, lr_G=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, D_L1=1e-7, D_L2=0.0, G_L1=0.0, G_L2=0.0,
             D_clamp=1.0, G_clamp=5.0)


def make_case(B, C, seed, init="trained"):
    """init: "trained" (slopes 0.25), "smooth" (slopes 1: no PReLU kinks, gradients strictly comparable; max-pool
    arg-max ties do not occur with continuous random inputs)."""
    rng = np.random.default_rng(seed)
    sl = 1.0 if init == "smooth" else 0.25
    gG, gD = (1.0, 0.8) if init == "smooth" else (1.2, 1.0)  # keeps D's outputs away from fp32 sigmoid saturation
    PG = LY.trained_like_init(LY.c2f_G_layout(C), rng, gG, slope=sl)
    PD = LY.trained_like_init(LY.c2f_D_layout(C), rng, gD, slope=sl)
    real_diff, cond_real = LY.c2f_pairs(B // 2, C, rng)
    _, cond_fake = LY.c2f_pairs(B // 2, C, rng)
    _, cond_G = LY.c2f_pairs(B, C, rng)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(PG=f(PG), PD=f(PD), real_diff=real_diff, cond_D=f(np.concatenate([cond_real, cond_fake])),
                noise_D=f(rng.uniform(-1, 1, (B // 2, 1, 32, 32))), cond_G=cond_G,
                noise_G=f(rng.uniform(-1, 1, (B, 1, 32, 32))),
                masks_D=f(rng.random((B, OC.MASK_PER_SAMPLE)) < 0.5), masks_G=f(rng.random((B, OC.MASK_PER_SAMPLE)) < 0.5))


def fresh_state(case, dtype=np.float64):
    PD, PG = case["PD"].astype(dtype), case["PG"].astype(dtype)
    return dict(PD=PD, PG=PG, mD=np.zeros_like(PD), vD=np.zeros_like(PD), mG=np.zeros_like(PG), vG=np.zeros_like(PG),
                tD=0, tG=0)


def oracle_iteration(case, B, C, hyper=None, state=None):
    st = state or fresh_state(case)
    res = OC.f64.train_iteration(B, C, hyper or HYPER, case["real_diff"], case["cond_D"], case["noise_D"], case["cond_G"],
                                 case["noise_G"], case["masks_D"], case["masks_G"], st)
    res["state"] = st
    return res
