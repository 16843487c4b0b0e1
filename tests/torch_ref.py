"""Independent (test-only) restatement of G / D / the train iteration on PyTorch-CPU autograd, float64.

Used ONLY to cross-check the CPU oracle (oracle/fg_oracle.cpp); PyTorch's THNN-lineage ops agree with
the Torch7 `nn` semantics except for BCE's eps form and SpatialDropout's missing rescale, which are
hand-coded here.  Never imported by the product.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle as O


def _split(P, layout):
    return {k: P[o:o + int(np.prod(s))].reshape(s) for k, (o, s) in layout.items()}


class SigmoidF32Out(torch.autograd.Function):
    """D's output sigmoid as the reference's criterion / backward see it: the value is a float32 (the reference
    computes D in fp32 and hands a FloatTensor to nn.BCECriterion, train.lua:148), and the
    backward uses that stored float32 output: dz = g * y * (1 - y).  Mirrors d_output() in oracle/fg_oracle.cpp."""

    @staticmethod
    def forward(ctx, z):
        y = torch.sigmoid(z).to(torch.float32).to(z.dtype)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return g * y * (1 - y)


def d_sigmoid(z):
    return SigmoidF32Out.apply(z)


def prelu(x, a, branch=None, name=None):
    """branch (tests at the headline batch size): callable(name, x) -> bool tensor "take the x > 0 branch" (lets a
    test force the branch of pre-activations that lie within rounding noise of 0, see parity_utils)."""
    pos = (x > 0) if branch is None else branch(name, x.detach())
    return torch.where(pos, x, a * x)


def G_forward(P, noise, C=3, branch=None):
    p = _split(P, O.G_layout(C))
    B = noise.shape[0]
    z0 = F.linear(noise, p["L1W"], p["L1b"]).view(B, 128, 8, 8)
    h0 = prelu(z0, p["a1"], branch, "z0")
    u0 = F.interpolate(h0, scale_factor=2, mode="nearest")
    z1 = F.conv2d(u0, p["C1W"], p["C1b"], padding=2)
    y1 = F.batch_norm(z1, None, None, p["g1"], p["be1"], training=True, momentum=0.1, eps=1e-5)
    h1 = prelu(y1, p["a2"], branch, "y1")
    u1 = F.interpolate(h1, scale_factor=2, mode="nearest")
    z2 = F.conv2d(u1, p["C2W"], p["C2b"], padding=2)
    y2 = F.batch_norm(z2, None, None, p["g2"], p["be2"], training=True, momentum=0.1, eps=1e-5)
    h2 = prelu(y2, p["a3"], branch, "y2")
    z3 = F.conv2d(h2, p["C3W"], p["C3b"], padding=1)
    return torch.sigmoid(z3), dict(z0=z0, h0=h0, z1=z1, y1=y1, h1=h1, z2=z2, y2=y2, h2=h2, z3=z3)


def D_forward(P, img, masks, C=3, branch=None):
    p = _split(P, O.D_layout(C))
    B = img.shape[0]
    moff = [0, 64, 192, 448]
    cout = [64, 128, 256, 512]
    x = img
    for i in range(4):
        z = F.conv2d(x, p["c%dW" % (i + 1)], p["c%db" % (i + 1)], padding=1)
        a = prelu(z, p["a%d" % (i + 1)], branch, "z%d" % (i + 1))
        m = masks[:, moff[i]:moff[i] + cout[i]].reshape(B, cout[i], 1, 1)
        x = F.avg_pool2d(a * m, 2, 2)  # SpatialDropout: no 1/(1-p) rescale in training
    x = x.reshape(B, 2048)
    h = prelu(F.linear(x, p["L1W"], p["L1b"]), p["a5"], branch, "zl1") * masks[:, 960:1472] * 2.0
    h = prelu(F.linear(h, p["L2W"], p["L2b"]), p["a6"], branch, "zl2") * masks[:, 1472:1984] * 2.0
    return d_sigmoid(F.linear(h, p["L3W"], p["L3b"])).reshape(B)


def bce(x, t):
    eps = 1e-12
    return -(t * torch.log(x + eps) + (1 - t) * torch.log(1 - x + eps)).mean()


def bce_grad(x, t):
    """The 2015 Lua nn.BCECriterion gradient: -(t-x)/(x(1-x+eps)+eps)/N  (NOT autograd of bce())."""
    eps = 1e-12
    return -(t - x) / (x * (1 - x + eps) + eps) / x.numel()


def init_params(n, rng, wstd=0.005):
    return rng.standard_normal(n) * wstd


def trained_like_G(C, rng):
    """Non-degenerate init: weights N(0,0.05^2), gamma U(0.5,1.5), slopes 0.25."""
    L = O.G_layout(C)
    P = np.zeros(O.G_param_count(C))
    for k, (o, s) in L.items():
        n = int(np.prod(s))
        if k in ("a1", "a2", "a3"):
            P[o] = 0.25
        elif k in ("g1", "g2"):
            P[o:o + n] = rng.uniform(0.5, 1.5, n)
        elif k.endswith("W"):
            fan_in = int(np.prod(s[1:]))
            P[o:o + n] = rng.standard_normal(n) * (1.0 / np.sqrt(fan_in))
        else:
            P[o:o + n] = rng.standard_normal(n) * 0.05
    return P


def trained_like_D(C, rng):
    L = O.D_layout(C)
    P = np.zeros(O.D_param_count(C))
    for k, (o, s) in L.items():
        n = int(np.prod(s))
        if k.startswith("a"):
            P[o] = 0.25
        elif k.endswith("W"):
            fan_in = int(np.prod(s[1:]))
            P[o:o + n] = rng.standard_normal(n) * (1.4 / np.sqrt(fan_in))
        else:
            P[o:o + n] = rng.standard_normal(n) * 0.05
    return P


def make_masks(B, rng):
    m = np.zeros((B, O.MASK_PER_SAMPLE))
    m[:, :960] = rng.random((B, 960)) < 0.8
    m[:, 960:] = rng.random((B, 1024)) < 0.5
    return m
