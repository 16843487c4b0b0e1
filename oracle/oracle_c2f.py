"""ctypes front-end of the coarse-to-fine part of oracle/libfg_oracle.so (fg_oracle_c2f.h).

TEST INFRASTRUCTURE ONLY (same rules as oracle.py).  PARITY UNPINNED.
Restates models_c2f.lua:113-145 (create_G_d), :237-278 (create_D_c) and adversarial_c2f.lua:121-187.
"""
import ctypes as C

import numpy as np

from . import oracle as O

MASK_PER_SAMPLE = 16384 + 512


def _lib():
    lib = O.lib()
    if not getattr(lib, "_c2f_ready", False):
        lib.fgo_c2f_G_param_count.restype = C.c_long
        lib.fgo_c2f_D_param_count.restype = C.c_long
        for sfx in ("f64", "f32"):
            getattr(lib, "fgo_c2f_G_new_" + sfx).restype = C.c_void_p
            getattr(lib, "fgo_c2f_D_new_" + sfx).restype = C.c_void_p
        lib._c2f_ready = True
    return lib


def G_param_count(c):
    return int(_lib().fgo_c2f_G_param_count(c))


def D_param_count(c):
    return int(_lib().fgo_c2f_D_param_count(c))


def G_layout(c):
    """name -> (offset, shape), getParameters order of create_G_d (models_c2f.lua:124-133)."""
    cin, cout, k = [c + 1, 64, 64, 128, 256], [64, 64, 128, 256, c], [3, 3, 5, 5, 7]
    out, o = {}, 0
    for i in range(5):
        items = [("c%dW" % (i + 1), (cout[i], cin[i], k[i], k[i])), ("c%db" % (i + 1), (cout[i],))]
        if i < 4:
            items.append(("a%d" % (i + 1), (1,)))
        for name, shape in items:
            out[name] = (o, shape)
            o += int(np.prod(shape))
    assert o == G_param_count(c)
    return out


def D_layout(c):
    """name -> (offset, shape), getParameters order of create_D_c (models_c2f.lua:247-265)."""
    cin, cout = [c, 64, 64, 128], [64, 64, 128, 256]
    items = []
    for i in range(4):
        items += [("c%dW" % (i + 1), (cout[i], cin[i], 3, 3)), ("c%db" % (i + 1), (cout[i],)), ("a%d" % (i + 1), (1,))]
    items += [("L1W", (512, 16384)), ("L1b", (512,)), ("a5", (1,)), ("L2W", (1, 512)), ("L2b", (1,))]
    out, o = {}, 0
    for name, shape in items:
        out[name] = (o, shape)
        o += int(np.prod(shape))
    assert o == D_param_count(c)
    return out


class _C2f:
    def __init__(self, t):
        self.t = t
        _lib()

    def maxpool2_fwd(self, x):
        t = self.t
        x = t.a(x)
        B, Cc, H, W = x.shape
        y = np.empty((B, Cc, H // 2, W // 2), t.dtype)
        arg = np.empty((B, Cc, H // 2, W // 2), np.uint8)
        t.f("maxpool2_fwd")(B * Cc, H, W, t.p(x), t.p(y), t.p(arg))
        return y, arg

    def maxpool2_bwd(self, dy, arg):
        t = self.t
        dy = t.a(dy)
        arg = np.ascontiguousarray(arg, np.uint8)
        B, Cc, Ho, Wo = dy.shape
        dx = np.empty((B, Cc, 2 * Ho, 2 * Wo), t.dtype)
        t.f("maxpool2_bwd")(B * Cc, 2 * Ho, 2 * Wo, t.p(dy), t.p(arg), t.p(dx))
        return dx

    def G(self):
        return _GNet(self.t)

    def D(self):
        return _DNet(self.t)

    def train_iteration(self, B, Cc, hyper, real_diff, condD, noiseD, condG, noiseG, masksD, masksG, state,
                        want_grads=True):
        """state: dict PD,PG,mD,vD,mG,vG (arrays of the oracle dtype, updated in place), tD,tG ints."""
        t = self.t
        hp = np.array([hyper[k] for k in ("lr_D", "lr_G", "beta1", "beta2", "eps", "D_L1", "D_L2", "G_L1", "G_L2",
                                          "D_clamp", "G_clamp")], np.float64)
        real_diff, condD, noiseD, condG, noiseG, masksD, masksG = map(
            t.a, (real_diff, condD, noiseD, condG, noiseG, masksD, masksG))
        tD, tG = C.c_int(state["tD"]), C.c_int(state["tG"])
        stats = np.zeros(8, np.float64)
        gD = np.zeros(state["PD"].size, t.dtype) if want_grads else None
        gG = np.zeros(state["PG"].size, t.dtype) if want_grads else None
        fake = np.zeros((B // 2, Cc, 32, 32), t.dtype)
        outD = np.zeros(B, t.dtype)
        t.f("c2f_train_iteration")(B, Cc, t.p(hp), t.p(real_diff), t.p(condD), t.p(noiseD), t.p(condG), t.p(noiseG),
                                   t.p(masksD), t.p(masksG), t.p(state["PD"]), t.p(state["PG"]), t.p(state["mD"]),
                                   t.p(state["vD"]), t.p(state["mG"]), t.p(state["vG"]), C.byref(tD), C.byref(tG),
                                   t.p(stats), t.p(gD), t.p(gG), t.p(fake), t.p(outD))
        state["tD"], state["tG"] = tD.value, tG.value
        return dict(lossD=stats[0], lossG=stats[1], conf=stats[2:6].copy(), gradD=gD, gradG=gG, fake=fake, outD=outD)


class _GNet:
    def __init__(self, t):
        self.t = t
        self.h = C.c_void_p(t.f("c2f_G_new")())

    def __del__(self):
        try:
            self.t.f("c2f_G_free")(self.h)
        except Exception:
            pass

    def forward(self, P, noise, cond):
        t = self.t
        self.P, noise, cond = t.a(P), t.a(noise), t.a(cond)
        B, Cc = cond.shape[0], cond.shape[1]
        self.B, self.C = B, Cc
        out = np.empty((B, Cc, 32, 32), t.dtype)
        t.f("c2f_G_forward")(self.h, t.p(self.P), t.p(noise), t.p(cond), B, Cc, t.p(out))
        return out

    def backward(self, dout):
        t = self.t
        dout = t.a(dout)
        dP = np.zeros(self.P.size, t.dtype)
        t.f("c2f_G_backward")(self.h, t.p(self.P), t.p(dout), t.p(dP))
        return dP


class _DNet:
    def __init__(self, t):
        self.t = t
        self.h = C.c_void_p(t.f("c2f_D_new")())

    def __del__(self):
        try:
            self.t.f("c2f_D_free")(self.h)
        except Exception:
            pass

    def forward(self, P, diff, cond, masks=None, training=True):
        t = self.t
        self.P, diff, cond = t.a(P), t.a(diff), t.a(cond)
        B, Cc = diff.shape[0], diff.shape[1]
        self.B, self.C = B, Cc
        masks = t.a(masks) if masks is not None else None
        out = np.empty(B, t.dtype)
        t.f("c2f_D_forward")(self.h, t.p(self.P), t.p(diff), t.p(cond), B, Cc, int(training), t.p(masks), t.p(out))
        return out

    def backward(self, dout, want_dP=True, want_ddiff=True):
        t = self.t
        dout = t.a(dout)
        dP = np.zeros(self.P.size, t.dtype) if want_dP else None
        dd = np.zeros((self.B, self.C, 32, 32), t.dtype) if want_ddiff else None
        t.f("c2f_D_backward")(self.h, t.p(self.P), t.p(dout), t.p(dP), t.p(dd))
        return dP, dd


f64 = _C2f(O.f64)
f32 = _C2f(O.f32)
