"""ctypes front-end of the --scale 16 part of oracle/libfg_oracle.so (fg_oracle_s16.h).
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.  models.lua:26-51 (create_G_decoder_upsampling16), :279-316 (create_D16_d)."""
import ctypes as C

import numpy as np

from . import oracle as O

MASK_PER_SAMPLE = 1024 + 128


def _lib():
    lib = O.lib()
    if not getattr(lib, "_s16_ready", False):
        lib.fgo_s16_G_param_count.restype = C.c_long
        lib.fgo_s16_D_param_count.restype = C.c_long
        for sfx in ("f64", "f32"):
            getattr(lib, "fgo_s16_G_new_" + sfx).restype = C.c_void_p
            getattr(lib, "fgo_s16_D_new_" + sfx).restype = C.c_void_p
        lib._s16_ready = True
    return lib


def G_param_count(c):
    return int(_lib().fgo_s16_G_param_count(c))


def D_param_count(c):
    return int(_lib().fgo_s16_D_param_count(c))


def _layout(items, count):
    out, o = {}, 0
    for name, shape in items:
        out[name] = (o, shape)
        o += int(np.prod(shape))
    assert o == count
    return out


def G_layout(c):
    return _layout([("L1W", (2048, 100)), ("L1b", (2048,)), ("a1", (1,)), ("C1W", (256, 128, 5, 5)), ("C1b", (256,)),
                    ("g1", (256,)), ("be1", (256,)), ("a2", (1,)), ("C2W", (128, 256, 5, 5)), ("C2b", (128,)), ("g2", (128,)),
                    ("be2", (128,)), ("a3", (1,)), ("C3W", (c, 128, 3, 3)), ("C3b", (c,))], G_param_count(c))


def D_layout(c):
    cin, cout = [c, 128, 128, 512], [128, 128, 512, 1024]
    items = []
    for i in range(4):
        items += [("c%dW" % (i + 1), (cout[i], cin[i], 3, 3)), ("c%db" % (i + 1), (cout[i],)), ("a%d" % (i + 1), (1,))]
    items += [("F1W", (1024, 4096)), ("F1b", (1024,)), ("af", (1,)), ("E1W", (128, c * 256)), ("E1b", (128,)), ("ae1", (1,)),
              ("E2W", (128, 128)), ("E2b", (128,)), ("ae2", (1,)), ("JW", (1, 1152)), ("Jb", (1,))]
    return _layout(items, D_param_count(c))


class _S16:
    def __init__(self, t):
        self.t = t
        _lib()

    def convs_fwd(self, x, W, b, stride, pad):
        t = self.t
        x, W, b = t.a(x), t.a(W), t.a(b)
        B, Cin, H, Wd = x.shape
        k = W.shape[2]
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (Wd + 2 * pad - k) // stride + 1
        y = np.empty((B, W.shape[0], Ho, Wo), t.dtype)
        t.f("convs_fwd")(B, Cin, H, Wd, W.shape[0], k, stride, pad, t.p(x), t.p(W), t.p(b), t.p(y))
        return y

    def convs_bwd(self, x, W, dy, stride, pad):
        t = self.t
        x, W, dy = t.a(x), t.a(W), t.a(dy)
        B, Cin, H, Wd = x.shape
        dx, dW, db = np.zeros_like(x), np.zeros_like(W), np.zeros(W.shape[0], t.dtype)
        t.f("convs_bwd")(B, Cin, H, Wd, W.shape[0], W.shape[2], stride, pad, t.p(x), t.p(W), t.p(dy), t.p(dx), t.p(dW), t.p(db))
        return dx, dW, db

    def G(self):
        return _GNet(self.t)

    def D(self):
        return _DNet(self.t)


class _GNet:
    def __init__(self, t):
        self.t = t
        self.h = C.c_void_p(t.f("s16_G_new")())

    def __del__(self):
        try:
            self.t.f("s16_G_free")(self.h)
        except Exception:
            pass

    def forward(self, P, noise, Cc=3, bn_state=None):
        t = self.t
        self.P, noise = t.a(P), t.a(noise)
        B = noise.shape[0]
        self.B, self.C = B, Cc
        out = np.empty((B, Cc, 16, 16), t.dtype)
        t.f("s16_G_forward")(self.h, t.p(self.P), t.p(noise), B, Cc, t.p(bn_state), t.p(out))
        return out

    def backward(self, dout):
        t = self.t
        dP = np.zeros(self.P.size, t.dtype)
        t.f("s16_G_backward")(self.h, t.p(self.P), t.p(t.a(dout)), t.p(dP))
        return dP


class _DNet:
    def __init__(self, t):
        self.t = t
        self.h = C.c_void_p(t.f("s16_D_new")())

    def __del__(self):
        try:
            self.t.f("s16_D_free")(self.h)
        except Exception:
            pass

    def forward(self, P, img, masks=None, training=True):
        t = self.t
        self.P, img = t.a(P), t.a(img)
        B, Cc = img.shape[0], img.shape[1]
        self.B, self.C = B, Cc
        masks = t.a(masks) if masks is not None else None
        out = np.empty(B, t.dtype)
        t.f("s16_D_forward")(self.h, t.p(self.P), t.p(img), B, Cc, int(training), t.p(masks), t.p(out))
        return out

    def backward(self, dout):
        t = self.t
        dP = np.zeros(self.P.size, t.dtype)
        dimg = np.zeros((self.B, self.C, 16, 16), t.dtype)
        t.f("s16_D_backward")(self.h, t.p(self.P), t.p(t.a(dout)), t.p(dP), t.p(dimg))
        return dP, dimg


f64 = _S16(O.f64)
f32 = _S16(O.f32)
