"""Shared helpers for the parity tests: seeded cases, oracle runs, error norms."""
import numpy as np

from oracle import oracle as O
from face_generator_b200 import layouts as LY

HYPER = dict(lr_D=1e-3, lr_G=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, D_L1=0.0, D_L2=1e-4, G_L1=0.0, G_L2=0.0,
             D_clamp=1.0, G_clamp=5.0)


def relerr(a, b):
    """normwise relative error max|a-b| / max|b| (the parity bar is 1e-4, BASELINE.json north_star)."""
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def make_masks(B, rng):
    m = np.zeros((B, O.MASK_PER_SAMPLE))
    m[:, :960] = rng.random((B, 960)) < 0.8
    m[:, 960:] = rng.random((B, 1024)) < 0.5
    return m


def make_case(B, C, seed, init="trained"):
    rng = np.random.default_rng(seed)
    if init in ("trained", "smooth"):
        # "smooth": all PReLU slopes = 1 (identity) -> the loss is differentiable everywhere, so gradients can be
        # compared strictly; with real slopes a pre-activation within fp32 noise of 0 may take the other branch
        sl = 1.0 if init == "smooth" else 0.25
        PG = LY.trained_like_init(LY.G_layout(C), rng, slope=sl)
        PD = LY.trained_like_init(LY.D_layout(C), rng, 1.4, slope=sl)
    else:  # the reference's own init: N(0, 0.005^2) weights (incl. BN gamma, PReLU slope), N(0, 0.001^2) biases
        PG, PD = LY.reference_init(LY.G_layout(C), rng), LY.reference_init(LY.D_layout(C), rng)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(PG=f(PG), PD=f(PD), real=f(rng.random((B // 2, C, 32, 32))),
                noise_D=f(rng.uniform(-1, 1, (B // 2, 100))), noise_G=f(rng.uniform(-1, 1, (B, 100))),
                masks_D=f(make_masks(B, rng)), masks_G=f(make_masks(B, rng)))


def fresh_state(case, dtype=np.float64):
    PD, PG = case["PD"].astype(dtype), case["PG"].astype(dtype)
    return dict(PD=PD, PG=PG, mD=np.zeros_like(PD), vD=np.zeros_like(PD), mG=np.zeros_like(PG), vG=np.zeros_like(PG),
                tD=0, tG=0, bnG=np.concatenate([np.zeros(256), np.ones(256), np.zeros(128), np.ones(128)]).astype(dtype))


def oracle_iteration(case, B, C, hyper=None, state=None):
    st = state or fresh_state(case)
    res = O.f64.train_iteration(B, C, hyper or HYPER, case["real"], case["noise_D"], case["noise_G"], case["masks_D"],
                                case["masks_G"], st)
    res["state"] = st
    return res
