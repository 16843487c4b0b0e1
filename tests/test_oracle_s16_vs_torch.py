"""CPU: the --scale 16 nets of the oracle (models.lua:26-51 create_G_decoder_upsampling16, :279-316 create_D16_d with
its stride-2 convolutions and ConcatTable/JoinTable) against PyTorch-CPU autograd in fp64.  No CUDA counterpart yet
(SURVEY.md 8(f).4): this is the checker it will be held to."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from torch_ref import d_sigmoid

from oracle import oracle_s16 as OS
from torch_ref import _split, prelu
from torch_ref_c2f import trained_like

torch.set_num_threads(8)


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-300)


@pytest.mark.parametrize("stride,pad,H,k", [(2, 1, 8, 3), (2, 1, 5, 3), (1, 1, 6, 3), (2, 2, 9, 5), (3, 0, 10, 3)])
def test_strided_conv_matches_torch(stride, pad, H, k):
    rng = np.random.default_rng(50 + stride + H)
    x, w, b = rng.standard_normal((2, 3, H, H + 1)), rng.standard_normal((4, 3, k, k)), rng.standard_normal(4)
    y = OS.f64.convs_fwd(x, w, b, stride, pad)
    xt, wt, bt = (torch.tensor(v, requires_grad=True) for v in (x, w, b))
    yt = F.conv2d(xt, wt, bt, stride=stride, padding=pad)
    assert y.shape == tuple(yt.shape) and rel(y, yt.detach().numpy()) < 1e-12
    dy = rng.standard_normal(y.shape)
    yt.backward(torch.tensor(dy))
    dx, dw, db = OS.f64.convs_bwd(x, w, dy, stride, pad)
    assert rel(dx, xt.grad.numpy()) < 1e-12 and rel(dw, wt.grad.numpy()) < 1e-12 and rel(db, bt.grad.numpy()) < 1e-12


def torch_G16(P, noise, C):
    p = _split(P, OS.G_layout(C))
    B = noise.shape[0]
    h = prelu(F.linear(noise, p["L1W"], p["L1b"]).view(B, 128, 4, 4), p["a1"])
    h = F.conv2d(F.interpolate(h, scale_factor=2, mode="nearest"), p["C1W"], p["C1b"], padding=2)
    h = prelu(F.batch_norm(h, None, None, p["g1"], p["be1"], training=True, eps=1e-5), p["a2"])
    h = F.conv2d(F.interpolate(h, scale_factor=2, mode="nearest"), p["C2W"], p["C2b"], padding=2)
    h = prelu(F.batch_norm(h, None, None, p["g2"], p["be2"], training=True, eps=1e-5), p["a3"])
    return torch.sigmoid(F.conv2d(h, p["C3W"], p["C3b"], padding=1))


def torch_D16(P, img, masks, C):
    p = _split(P, OS.D_layout(C))
    B = img.shape[0]
    h = prelu(F.conv2d(img, p["c1W"], p["c1b"], padding=1), p["a1"])
    h = prelu(F.conv2d(h, p["c2W"], p["c2b"], padding=1), p["a2"])
    h = F.avg_pool2d(h, 2, 2)
    h = prelu(F.conv2d(h, p["c3W"], p["c3b"], stride=2, padding=1), p["a3"])
    h = prelu(F.conv2d(h, p["c4W"], p["c4b"], stride=2, padding=1), p["a4"])
    h = h * masks[:, :1024].reshape(B, 1024, 1, 1)  # SpatialDropout: no rescale
    fine = prelu(F.linear(h.reshape(B, 4096), p["F1W"], p["F1b"]), p["af"])
    e = prelu(F.linear(img.reshape(B, -1), p["E1W"], p["E1b"]), p["ae1"]) * masks[:, 1024:] * 2.0
    e = prelu(F.linear(e, p["E2W"], p["E2b"]), p["ae2"])
    return d_sigmoid(F.linear(torch.cat([fine, e], dim=1), p["JW"], p["Jb"])).reshape(B)


def test_param_counts():
    # G16: the 32px generator minus 3/4 of its first Linear; D16_d from the layer list of models.lua:279-316
    assert OS.G_param_count(3) == 2470406 - 6144 * 101
    assert OS.D_param_count(3) == sum(int(np.prod(s)) for _, s in OS.D_layout(3).values())


@pytest.mark.parametrize("C", [3, 1])
def test_G16_fwd_bwd_matches_torch(C):
    rng = np.random.default_rng(60 + C)
    B = 4
    P = trained_like(OS.G_layout(C), OS.G_param_count(C), rng, gain=1.0)
    for k in ("g1", "g2"):
        o, s = OS.G_layout(C)[k]
        P[o:o + s[0]] = rng.uniform(0.5, 1.5, s[0])
    noise = rng.uniform(-1, 1, (B, 100))
    dout = rng.standard_normal((B, C, 16, 16))
    g = OS.f64.G()
    out = g.forward(P, noise, C)
    dP = g.backward(dout)
    Pt = torch.tensor(P, requires_grad=True)
    out_t = torch_G16(Pt, torch.tensor(noise), C)
    out_t.backward(torch.tensor(dout))
    assert rel(out, out_t.detach().numpy()) < 1e-11
    gt = Pt.grad.numpy()
    for k, (o, s) in OS.G_layout(C).items():
        n = int(np.prod(s))
        if k in ("C1b", "C2b"):  # bias in front of BatchNorm: analytically zero, rounding noise only
            assert np.abs(dP[o:o + n]).max() < 1e-9 * np.abs(gt).max()
            continue
        assert rel(dP[o:o + n], gt[o:o + n]) < 1e-8, k


@pytest.mark.parametrize("C", [3, 1])
def test_D16_fwd_bwd_matches_torch(C):
    rng = np.random.default_rng(70 + C)
    B = 3
    P = trained_like(OS.D_layout(C), OS.D_param_count(C), rng, gain=1.0)
    img = rng.random((B, C, 16, 16))
    masks = (rng.random((B, OS.MASK_PER_SAMPLE)) < 0.5).astype(np.float64)
    dout = rng.standard_normal(B)
    d = OS.f64.D()
    out = d.forward(P, img, masks)
    dP, dimg = d.backward(dout)
    Pt = torch.tensor(P, requires_grad=True)
    it = torch.tensor(img, requires_grad=True)
    out_t = torch_D16(Pt, it, torch.tensor(masks), C)
    out_t.backward(torch.tensor(dout))
    assert rel(out, out_t.detach().numpy()) < 1e-12
    gt = Pt.grad.numpy()
    for k, (o, s) in OS.D_layout(C).items():
        n = int(np.prod(s))
        assert rel(dP[o:o + n], gt[o:o + n]) < 1e-9, k
    assert rel(dimg, it.grad.numpy()) < 1e-9  # ConcatTable: both branches contribute to the input gradient
