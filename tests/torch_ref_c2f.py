"""Independent (test-only) PyTorch-CPU fp64 restatement of the coarse-to-fine nets
(models_c2f.lua:113-145 create_G_d, :237-278 create_D_c).  Cross-checks oracle/fg_oracle_c2f.h only."""
import numpy as np
import torch
import torch.nn.functional as F

from torch_ref import d_sigmoid

from oracle import oracle_c2f as OC
from torch_ref import _split, prelu


def G_forward(P, noise, cond, C=3):
    p = _split(P, OC.G_layout(C))
    x = torch.cat([noise, cond], dim=1)  # JoinTable(2,2): noise plane first
    pads = [1, 1, 2, 2, 3]
    for i in range(5):
        x = F.conv2d(x, p["c%dW" % (i + 1)], p["c%db" % (i + 1)], padding=pads[i])
        if i < 4:
            x = prelu(x, p["a%d" % (i + 1)])
    return x


def D_forward(P, diff, cond, masks, C=3):
    p = _split(P, OC.D_layout(C))
    B = diff.shape[0]
    x = diff + cond  # CAddTable
    for i in range(4):
        x = prelu(F.conv2d(x, p["c%dW" % (i + 1)], p["c%db" % (i + 1)], padding=1), p["a%d" % (i + 1)])
        if i in (1, 3):
            x = F.max_pool2d(x, 2, 2)
    x = x.reshape(B, 16384) * masks[:, :16384] * 2.0  # nn.Dropout p=0.5 (v2), then View in (c,h,w) order
    h = prelu(F.linear(x, p["L1W"], p["L1b"]), p["a5"]) * masks[:, 16384:] * 2.0
    return d_sigmoid(F.linear(h, p["L2W"], p["L2b"])).reshape(B)


def trained_like(layout, count, rng, gain=1.4):
    """He-style weights (activations stay O(1) through the PReLU stacks), slopes 0.25, small biases."""
    P = np.zeros(count)
    for k, (o, s) in layout.items():
        n = int(np.prod(s))
        if k.startswith("a"):
            P[o] = 0.25
        elif k.endswith("W"):
            P[o:o + n] = rng.standard_normal(n) * (gain / np.sqrt(int(np.prod(s[1:]))))
        else:
            P[o:o + n] = rng.standard_normal(n) * 0.05
    return P


def trained_like_G(C, rng):
    return trained_like(OC.G_layout(C), OC.G_param_count(C), rng)


def trained_like_D(C, rng):
    return trained_like(OC.D_layout(C), OC.D_param_count(C), rng)


def make_masks(B, rng):
    return (rng.random((B, OC.MASK_PER_SAMPLE)) < 0.5).astype(np.float64)


def make_pairs(B, C, rng):
    """Stand-in for dataset_c2f.lua:54-60: fine ~ U[0,1), coarse = 2x avg-down then 2x nearest-up, diff = fine - coarse."""
    fine = rng.random((B, C, 32, 32))
    small = fine.reshape(B, C, 16, 2, 16, 2).mean(axis=(3, 5))
    coarse = np.repeat(np.repeat(small, 2, axis=2), 2, axis=3)
    return fine - coarse, coarse
