"""ctypes front-end of oracle/libfg_oracle.so (the CPU restatement of the reference hot path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (face_generator_b200) never imports this.
PARITY UNPINNED -- see the header of fg_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfg_oracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("fg_oracle.cpp", "fg_oracle_c2f.h", "fg_oracle_s16.h")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.fgo_G_param_count.restype = C.c_long
        _lib.fgo_D_param_count.restype = C.c_long
        for sfx in ("f64", "f32"):
            getattr(_lib, "fgo_bce_fwd_" + sfx).restype = C.c_double
            getattr(_lib, "fgo_penalty_clamp_" + sfx).restype = C.c_double
            getattr(_lib, "fgo_G_new_" + sfx).restype = C.c_void_p
            getattr(_lib, "fgo_D_new_" + sfx).restype = C.c_void_p
            getattr(_lib, "fgo_G_tap_" + sfx).restype = C.c_long
    return _lib


MASK_PER_SAMPLE = 1984


def use_blas(threads=0):
    """Route the fp32 port's GEMMs through the OpenBLAS bundled with scipy (what Torch7's nn would call on a CPU
    box).  Returns the library path used, or None when no OpenBLAS is found (the blocked loops stay in use).
    Only the *_f32 entry points are affected; the fp64 parity oracle never uses BLAS."""
    import glob
    cands = []
    try:
        import scipy
        cands += sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs",
                                               "libscipy_openblas*.so")))
    except Exception:
        pass
    for path in cands:
        if lib().fgo_use_blas(path.encode(), int(threads)) == 1:
            return path
    lib().fgo_use_blas(None, 0)
    return None


def blas_active():
    return bool(lib().fgo_blas_active())


def G_param_count(c):
    return int(lib().fgo_G_param_count(c))


def D_param_count(c):
    return int(lib().fgo_D_param_count(c))


def G_layout(c):
    """name -> (offset, shape) of the flat G parameter vector (getParameters order; models.lua:57-81)."""
    out, o = {}, 0
    for name, shape in [("L1W", (8192, 100)), ("L1b", (8192,)), ("a1", (1,)), ("C1W", (256, 128, 5, 5)),
                        ("C1b", (256,)), ("g1", (256,)), ("be1", (256,)), ("a2", (1,)), ("C2W", (128, 256, 5, 5)),
                        ("C2b", (128,)), ("g2", (128,)), ("be2", (128,)), ("a3", (1,)), ("C3W", (c, 128, 3, 3)),
                        ("C3b", (c,))]:
        out[name] = (o, shape)
        o += int(np.prod(shape))
    assert o == G_param_count(c)
    return out


def D_layout(c):
    """name -> (offset, shape) of the flat D parameter vector (models.lua:382-416)."""
    out, o = {}, 0
    cin, cout = [c, 64, 128, 256], [64, 128, 256, 512]
    items = []
    for i in range(4):
        items += [("c%dW" % (i + 1), (cout[i], cin[i], 3, 3)), ("c%db" % (i + 1), (cout[i],)), ("a%d" % (i + 1), (1,))]
    items += [("L1W", (512, 2048)), ("L1b", (512,)), ("a5", (1,)), ("L2W", (512, 512)), ("L2b", (512,)),
              ("a6", (1,)), ("L3W", (1, 512)), ("L3b", (1,))]
    for name, shape in items:
        out[name] = (o, shape)
        o += int(np.prod(shape))
    assert o == D_param_count(c)
    return out


class _T:
    """dtype-specific view of the exported entry points."""

    def __init__(self, sfx, dtype):
        self.sfx, self.dtype = sfx, np.dtype(dtype)
        self.ct = C.c_double if sfx == "f64" else C.c_float

    def f(self, name):
        return getattr(lib(), "fgo_%s_%s" % (name, self.sfx))

    def a(self, x):
        return np.ascontiguousarray(x, dtype=self.dtype)

    @staticmethod
    def p(x):
        return None if x is None else x.ctypes.data_as(C.c_void_p)

    # ---- ops (NCHW) ----
    def linear_fwd(self, x, W, b):
        x, W, b = self.a(x), self.a(W), self.a(b)
        y = np.empty((x.shape[0], W.shape[0]), self.dtype)
        self.f("linear_fwd")(x.shape[0], W.shape[1], W.shape[0], self.p(x), self.p(W), self.p(b), self.p(y))
        return y

    def linear_bwd(self, x, W, dy):
        x, W, dy = self.a(x), self.a(W), self.a(dy)
        dx, dW, db = np.zeros_like(x), np.zeros_like(W), np.zeros(W.shape[0], self.dtype)
        self.f("linear_bwd")(x.shape[0], W.shape[1], W.shape[0], self.p(x), self.p(W), self.p(dy), self.p(dx),
                             self.p(dW), self.p(db))
        return dx, dW, db

    def conv_fwd(self, x, W, b):
        x, W, b = self.a(x), self.a(W), self.a(b)
        B, Cin, H, Wd = x.shape
        y = np.empty((B, W.shape[0], H, Wd), self.dtype)
        self.f("conv_fwd")(B, Cin, H, Wd, W.shape[0], W.shape[2], self.p(x), self.p(W), self.p(b), self.p(y))
        return y

    def conv_bwd(self, x, W, dy):
        x, W, dy = self.a(x), self.a(W), self.a(dy)
        B, Cin, H, Wd = x.shape
        dx, dW, db = np.zeros_like(x), np.zeros_like(W), np.zeros(W.shape[0], self.dtype)
        self.f("conv_bwd")(B, Cin, H, Wd, W.shape[0], W.shape[2], self.p(x), self.p(W), self.p(dy), self.p(dx),
                           self.p(dW), self.p(db))
        return dx, dW, db

    def up2_fwd(self, x):
        x = self.a(x)
        B, Cc, H, W = x.shape
        y = np.empty((B, Cc, 2 * H, 2 * W), self.dtype)
        self.f("up2_fwd")(B, Cc, H, W, self.p(x), self.p(y))
        return y

    def up2_bwd(self, dy):
        dy = self.a(dy)
        B, Cc, H2, W2 = dy.shape
        dx = np.empty((B, Cc, H2 // 2, W2 // 2), self.dtype)
        self.f("up2_bwd")(B, Cc, H2 // 2, W2 // 2, self.p(dy), self.p(dx))
        return dx

    def bn_fwd_train(self, x, g, be, rm=None, rv=None):
        x, g, be = self.a(x), self.a(g), self.a(be)
        B, Cc, H, W = x.shape
        y = np.empty_like(x)
        mean, istd = np.empty(Cc, self.dtype), np.empty(Cc, self.dtype)
        self.f("bn_fwd_train")(B, Cc, H * W, self.p(x), self.p(g), self.p(be), self.p(y), self.p(mean),
                               self.p(istd), self.p(rm), self.p(rv))
        return y, mean, istd

    def bn_bwd(self, x, g, mean, istd, dy):
        x, g, mean, istd, dy = map(self.a, (x, g, mean, istd, dy))
        B, Cc, H, W = x.shape
        dx, dg, db = np.empty_like(x), np.zeros(Cc, self.dtype), np.zeros(Cc, self.dtype)
        self.f("bn_bwd")(B, Cc, H * W, self.p(x), self.p(g), self.p(mean), self.p(istd), self.p(dy), self.p(dx),
                         self.p(dg), self.p(db))
        return dx, dg, db

    def prelu_fwd(self, x, a):
        x = self.a(x)
        y = np.empty_like(x)
        self.f("prelu_fwd")(C.c_long(x.size), self.p(x), self.ct(a), self.p(y))
        return y

    def prelu_bwd(self, x, a, dy):
        x, dy = self.a(x), self.a(dy)
        dx, da = np.empty_like(x), np.zeros(1, self.dtype)
        self.f("prelu_bwd")(C.c_long(x.size), self.p(x), self.ct(a), self.p(dy), self.p(dx), self.p(da))
        return dx, da[0]

    def avgpool2_fwd(self, x):
        x = self.a(x)
        B, Cc, H, W = x.shape
        y = np.empty((B, Cc, H // 2, W // 2), self.dtype)
        self.f("avgpool2_fwd")(B * Cc, H, W, self.p(x), self.p(y))
        return y

    def avgpool2_bwd(self, dy):
        dy = self.a(dy)
        B, Cc, Ho, Wo = dy.shape
        dx = np.empty((B, Cc, 2 * Ho, 2 * Wo), self.dtype)
        self.f("avgpool2_bwd")(B * Cc, 2 * Ho, 2 * Wo, self.p(dy), self.p(dx))
        return dx

    def bce_fwd(self, x, t):
        x, t = self.a(x).ravel(), self.a(t).ravel()
        return float(self.f("bce_fwd")(x.size, self.p(x), self.p(t)))

    def bce_bwd(self, x, t):
        x, t = self.a(x).ravel(), self.a(t).ravel()
        dx = np.empty_like(x)
        self.f("bce_bwd")(x.size, self.p(x), self.p(t), self.p(dx))
        return dx

    def penalty_clamp(self, p, g, l1_loss, l1_grad, l2, clampv):
        """in-place on g; returns the loss term (adversarial.lua:103-109,121-123)."""
        assert g.dtype == self.dtype and g.flags.c_contiguous
        p = self.a(p)
        return float(self.f("penalty_clamp")(C.c_long(p.size), self.p(p), self.p(g), C.c_double(l1_loss),
                                             C.c_double(l1_grad), C.c_double(l2), C.c_double(clampv)))

    def adam(self, x, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        """in-place on x, m, v (interruptable_optimizers.lua:49-94); t = step count after increment."""
        for arr in (x, m, v):
            assert arr.dtype == self.dtype and arr.flags.c_contiguous
        g = self.a(g)
        self.f("adam")(C.c_long(x.size), self.p(x), self.p(g), self.p(m), self.p(v), int(t), C.c_double(lr),
                       C.c_double(b1), C.c_double(b2), C.c_double(eps))

    # ---- whole nets ----
    def G(self):
        return _GNet(self)

    def D(self):
        return _DNet(self)

    def train_iteration(self, B, Cc, hyper, real, noiseD, noiseG, masksD, masksG, state, want_grads=True):
        """state: dict PD,PG,mD,vD,mG,vG (arrays of self.dtype, updated in place), tD,tG ints, bnG[768]."""
        hp = np.array([hyper[k] for k in ("lr_D", "lr_G", "beta1", "beta2", "eps", "D_L1", "D_L2", "G_L1", "G_L2",
                                          "D_clamp", "G_clamp")], np.float64)
        real, noiseD, noiseG, masksD, masksG = map(self.a, (real, noiseD, noiseG, masksD, masksG))
        tD, tG = C.c_int(state["tD"]), C.c_int(state["tG"])
        stats = np.zeros(8, np.float64)
        gD = np.zeros(state["PD"].size, self.dtype) if want_grads else None
        gG = np.zeros(state["PG"].size, self.dtype) if want_grads else None
        fake = np.zeros((B // 2, Cc, 32, 32), self.dtype)
        outD = np.zeros(B, self.dtype)
        self.f("train_iteration")(B, Cc, self.p(hp), self.p(real), self.p(noiseD), self.p(noiseG), self.p(masksD),
                                  self.p(masksG), self.p(state["PD"]), self.p(state["PG"]), self.p(state["mD"]),
                                  self.p(state["vD"]), self.p(state["mG"]), self.p(state["vG"]), C.byref(tD),
                                  C.byref(tG), self.p(state["bnG"]), self.p(stats), self.p(gD), self.p(gG),
                                  self.p(fake), self.p(outD))
        state["tD"], state["tG"] = tD.value, tG.value
        return dict(lossD=stats[0], lossG=stats[1], conf=stats[2:6].copy(), gradD=gD, gradG=gG, fake=fake, outD=outD)


class _GNet:
    TAPS = {"z0": 0, "h0": 1, "z1": 2, "h1": 3, "z2": 4, "h2": 5, "z3": 6}

    def __init__(self, t):
        self.t = t
        self.h = C.c_void_p(t.f("G_new")())

    def __del__(self):
        try:
            self.t.f("G_free")(self.h)
        except Exception:
            pass

    def forward(self, P, noise, Cc=3, training=True, bn_state=None):
        t = self.t
        self.P, noise = t.a(P), t.a(noise)
        B = noise.shape[0]
        self.B, self.C = B, Cc
        out = np.empty((B, Cc, 32, 32), t.dtype)
        t.f("G_forward")(self.h, t.p(self.P), t.p(noise), B, Cc, int(training), t.p(bn_state), t.p(out))
        return out

    def backward(self, dout, want_dnoise=False):
        t = self.t
        dout = t.a(dout)
        dP = np.zeros(self.P.size, t.dtype)
        dn = np.zeros((self.B, 100), t.dtype) if want_dnoise else None
        t.f("G_backward")(self.h, t.p(self.P), t.p(dout), t.p(dP), t.p(dn))
        return (dP, dn) if want_dnoise else dP

    def tap(self, name):
        t = self.t
        n = t.f("G_tap")(self.h, self.TAPS[name], None)
        dst = np.empty(n, t.dtype)
        t.f("G_tap")(self.h, self.TAPS[name], t.p(dst))
        return dst


class _DNet:
    def __init__(self, t):
        self.t = t
        self.h = C.c_void_p(t.f("D_new")())

    def __del__(self):
        try:
            self.t.f("D_free")(self.h)
        except Exception:
            pass

    def forward(self, P, img, masks=None, training=True):
        t = self.t
        self.P, img = t.a(P), t.a(img)
        B, Cc = img.shape[0], img.shape[1]
        self.B, self.C = B, Cc
        masks = t.a(masks) if masks is not None else None
        out = np.empty(B, t.dtype)
        t.f("D_forward")(self.h, t.p(self.P), t.p(img), B, Cc, int(training), t.p(masks), t.p(out))
        return out

    def backward(self, dout, want_dP=True, want_dimg=True):
        t = self.t
        dout = t.a(dout)
        dP = np.zeros(self.P.size, t.dtype) if want_dP else None
        dimg = np.zeros((self.B, self.C, 32, 32), t.dtype) if want_dimg else None
        t.f("D_backward")(self.h, t.p(self.P), t.p(dout), t.p(dP), t.p(dimg))
        return dP, dimg


f64 = _T("f64", np.float64)
f32 = _T("f32", np.float32)


def num_threads():
    return int(lib().fgo_num_threads())


def set_num_threads(n):
    lib().fgo_set_num_threads(int(n))


# ---- PReLU kink bookkeeping (fg_oracle.cpp KinkCtx; used by the strict gradient-parity tests) -------------------
class kink:
    """with kink.record(margin): run the oracle -> kink.calls() lists, per prelu_fwd call (in call order), the
    pre-activation elements with |x| < margin*max|x|.  kink.set_override(call, idx, positive) + with kink.override():
    re-run -> prelu_bwd takes the given branch for those elements.  Call numbering restarts on every `with`."""

    class _Mode:
        def __init__(self, mode, margin):
            self.mode, self.margin = mode, margin

        def __enter__(self):
            lib().fgo_kink_mode(self.mode, C.c_double(self.margin))
            return self

        def __exit__(self, *a):
            lib().fgo_kink_mode(0, C.c_double(0.0))

    @staticmethod
    def record(margin):
        lib().fgo_kink_clear()
        return kink._Mode(1, margin)

    @staticmethod
    def override():
        return kink._Mode(2, 0.0)

    @staticmethod
    def clear():
        lib().fgo_kink_clear()

    @staticmethod
    def calls(ncalls):
        """[(n, maxabs, idx int64 array)] for calls 0..ncalls-1 (read after the `with kink.record()` block)."""
        L = lib()
        L.fgo_kink_call_info.restype = C.c_long
        out = []
        for s in range(ncalls):
            n, mx = C.c_long(0), C.c_double(0)
            cnt = L.fgo_kink_call_info(s, C.byref(n), C.byref(mx))
            assert cnt >= 0, "prelu_fwd call %d was not recorded" % s
            idx = np.empty(cnt, np.int64)
            if cnt:
                L.fgo_kink_call_indices(s, idx.ctypes.data_as(C.c_void_p))
            out.append((int(n.value), float(mx.value), idx))
        return out

    @staticmethod
    def set_override(call, idx, positive):
        idx = np.ascontiguousarray(idx, np.int64)
        pos = np.ascontiguousarray(positive, np.int8)
        assert idx.size == pos.size
        lib().fgo_kink_set_override(int(call), C.c_long(idx.size), idx.ctypes.data_as(C.c_void_p),
                                    pos.ctypes.data_as(C.c_void_p))
