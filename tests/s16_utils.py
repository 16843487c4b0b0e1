"""Seeded cases and the composed fp64 reference iteration for the --scale 16 nets (models.lua:27-51, :279-316 on the
adversarial.lua loop).  Test infrastructure: built from oracle/ pieces only."""
import numpy as np

from face_generator_b200 import layouts as LY
from oracle import oracle as O
from oracle import oracle_s16 as OS

# train.lua defaults (same loop as the 32x32 nets)
HYPER = dict(lr_D=1e-3, lr_G=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, D_L1=0.0, D_L2=1e-4, G_L1=0.0, G_L2=0.0,
             D_clamp=1.0, G_clamp=5.0)


def _slopes(P, layout, init):
    """"smooth": every PReLU slope 1 (no kinks).  "near": slopes 1 - k*1e-3, distinct per layer: the derivative jump
    at 0 is ~1e-3, so the handful of pre-activations that two correct fp32 implementations may sign differently move
    the batch-summed gradients by ~1e-6 (far below the 1e-4 bar), while a wrong x<=0 branch shows at 1e-3 in the
    forward values and the slope gradients sum_{x<=0} dh*x are compared at full scale.  "trained": 0.25."""
    k = 0
    for name, (o, s) in layout.items():
        if name[0] == "a":
            k += 1
            P[o] = 1.0 if init == "smooth" else (1.0 - 1e-3 * k if init == "near" else 0.25)
    return P


def make_case(B, C, seed, init="near"):
    rng = np.random.default_rng(seed)
    LG, LD = OS.G_layout(C), OS.D_layout(C)
    PG = _slopes(LY.trained_like_init((LG, OS.G_param_count(C)), rng, 1.0), LG, init)
    PD = _slopes(LY.trained_like_init((LD, OS.D_param_count(C)), rng, 0.8), LD, init)
    o, s = LD["JW"]
    PD[o:o + 1152] *= 1.0  # keeps D's outputs away from fp32 sigmoid saturation
    f = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(PG=f(PG), PD=f(PD), real=f(rng.random((B // 2, C, 16, 16))), noise_D=f(rng.uniform(-1, 1, (B // 2, 100))),
                noise_G=f(rng.uniform(-1, 1, (B, 100))), masks_D=f(rng.random((B, OS.MASK_PER_SAMPLE)) < 0.5),
                masks_G=f(rng.random((B, OS.MASK_PER_SAMPLE)) < 0.5))


def bn_init():
    s = np.zeros(768)
    s[256:512] = 1.0
    s[640:] = 1.0
    return s


def oracle_iteration(case, B, C, hyper=None):
    """One adversarial.lua:240-288 iteration (gate open) composed like oracle/fg_oracle.cpp train_iteration, on the 16x16 nets."""
    hp = hyper or HYPER
    t = O.f64
    PD, PG = case["PD"].astype(np.float64), case["PG"].astype(np.float64)
    mD, vD, mG, vG = np.zeros_like(PD), np.zeros_like(PD), np.zeros_like(PG), np.zeros_like(PG)
    bn = bn_init()
    Bh = B // 2
    g, d = OS.f64.G(), OS.f64.D()
    # ---- D step ----
    fake = g.forward(PG, case["noise_D"], C, bn)
    x = np.concatenate([case["real"].astype(np.float64), fake])
    tg = np.concatenate([np.ones(Bh), np.zeros(Bh)])
    outD = d.forward(PD, x, case["masks_D"], True)
    lossD = t.bce_fwd(outD, tg)
    gD, _ = d.backward(t.bce_bwd(outD, tg))
    lossD += t.penalty_clamp(PD, gD, hp["D_L1"], hp["D_L1"], hp["D_L2"], hp["D_clamp"])
    conf = np.zeros(4)
    for i in range(B):
        conf[(0 if outD[i] > 0.5 else 1) + (0 if i < Bh else 2)] += 1
    gradD = gD.copy()
    t.adam(PD, gD, mD, vD, 1, hp["lr_D"], hp["beta1"], hp["beta2"], hp["eps"])
    # ---- G step ----
    img = g.forward(PG, case["noise_G"], C, bn)
    outG = d.forward(PD, img, case["masks_G"], True)
    ones = np.ones(B)
    lossG = t.bce_fwd(outG, ones)
    _, dimg = d.backward(t.bce_bwd(outG, ones))
    gG = g.backward(dimg)
    lossG += t.penalty_clamp(PG, gG, hp["G_L1"], hp["G_L2"], hp["G_L2"], hp["G_clamp"])
    gradG = gG.copy()
    t.adam(PG, gG, mG, vG, 1, hp["lr_G"], hp["beta1"], hp["beta2"], hp["eps"])
    return dict(lossD=lossD, lossG=lossG, conf=conf, gradD=gradD, gradG=gradG, fake=fake, outD=outD, outG=outG, PD=PD, PG=PG,
                mD=mD, vD=vD, mG=mG, vG=vG, bn=bn)
