"""Torch7 checkpoint files (`torch.save` binary format) through the C ABI's fg_t7_* entry points.

Reading: the reference's `adversarial.net` = {D = MODEL_D, G = MODEL_G, opt = OPT, epoch = EPOCH}
(adversarial.lua:328, adversarial_c2f.lua:216; consumed by sample.lua:251-258 and train.lua:104-124).
Writing: flat float tensors + numbers + strings in one root table, readable with stock `torch.load`
(parameters, Adam moments and step counters -- the reference itself drops its optimizer state, train.lua:122).
Host-side only: nothing here touches the GPU.
"""
import ctypes as C

import numpy as np

from .lib import FGError, NET_D, NET_G, load_library

KINDS = {0: "nil", 1: "number", 2: "string", 3: "table", 4: "object", 5: "boolean", 6: "function", 16: "tensor",
         17: "storage", -1: None}


def _err(what):
    raise FGError("%s: %s" % (what, load_library().fg_last_error().decode()))


class T7File:
    def __init__(self, path):
        self.lib = load_library()
        h = C.c_void_p()
        if self.lib.fg_t7_open(str(path).encode(), C.byref(h)) != 0:
            _err("fg_t7_open")
        self.h = h

    def close(self):
        if self.h:
            self.lib.fg_t7_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def kind(self, path):
        return KINDS.get(int(self.lib.fg_t7_kind(self.h, path.encode())), "?")

    def number(self, path):
        v = C.c_double(0)
        if self.lib.fg_t7_number(self.h, path.encode(), C.byref(v)) != 0:
            _err("fg_t7_number(%s)" % path)
        return v.value

    def string(self, path):
        n = int(self.lib.fg_t7_string(self.h, path.encode(), None, 0))
        if n < 0:
            raise FGError("fg_t7_string(%s): not a string / object" % path)
        buf = C.create_string_buffer(n + 1)
        self.lib.fg_t7_string(self.h, path.encode(), buf, n + 1)
        return buf.value.decode()

    def tensor(self, path):
        dims = (C.c_int64 * 8)()
        n = int(self.lib.fg_t7_tensor(self.h, path.encode(), None, 0, dims))
        if n < 0:
            _err("fg_t7_tensor(%s)" % path)
        shape = [int(d) for d in dims if d > 0]
        out = np.empty(n, np.float32)
        if n and self.lib.fg_t7_tensor(self.h, path.encode(), out.ctypes.data_as(C.c_void_p), n, None) < 0:
            _err("fg_t7_tensor(%s)" % path)
        return out.reshape(shape) if n else out

    def _vec(self, fn, path):
        n = int(fn(self.h, path.encode(), None, 0))
        if n < 0:
            _err("%s(%s)" % (fn.__name__, path))
        out = np.empty(n, np.float32)
        if n and fn(self.h, path.encode(), out.ctypes.data_as(C.c_void_p), n) < 0:
            _err("%s(%s)" % (fn.__name__, path))
        return out

    def net_params(self, path):
        """Flat parameter vector of the module tree at `path` in getParameters() order."""
        return self._vec(self.lib.fg_t7_net_params, path)

    def net_bn_state(self, path):
        return self._vec(self.lib.fg_t7_net_bn_state, path)

    def net_describe(self, path):
        n = int(self.lib.fg_t7_net_describe(self.h, path.encode(), None, 0))
        if n < 0:
            raise FGError("fg_t7_net_describe(%s): no such entry" % path)
        buf = C.create_string_buffer(n + 1)
        self.lib.fg_t7_net_describe(self.h, path.encode(), buf, n + 1)
        return buf.value.decode()


class T7Writer:
    def __init__(self, path):
        self.lib = load_library()
        h = C.c_void_p()
        if self.lib.fg_t7_writer_open(str(path).encode(), C.byref(h)) != 0:
            _err("fg_t7_writer_open")
        self.h = h

    def add(self, key, value):
        k = key.encode()
        if isinstance(value, str):
            rc = self.lib.fg_t7_writer_add_string(self.h, k, value.encode())
        elif isinstance(value, (int, float, np.integer, np.floating)):
            rc = self.lib.fg_t7_writer_add_number(self.h, k, float(value))
        else:
            a = np.ascontiguousarray(value, np.float32)
            dims = (C.c_int64 * a.ndim)(*a.shape)
            rc = self.lib.fg_t7_writer_add_tensor(self.h, k, a.ctypes.data_as(C.c_void_p), dims, a.ndim)
        if rc != 0:
            _err("fg_t7_writer_add(%s)" % key)

    def close(self):
        if self.h:
            rc = self.lib.fg_t7_writer_close(self.h)
            self.h = None
            if rc != 0:
                _err("fg_t7_writer_close")


def load_reference_checkpoint(ctx, path, want_D=True):
    """sample.lua:247-258 `loadModels`: upload G (and D) of a reference `adversarial.net` into a Context.
    Raises if the stored nets are not the default 32x32 architectures this library implements."""
    with T7File(path) as f:
        pg = f.net_params("G")
        if pg.size != ctx.count(NET_G):
            raise FGError("checkpoint G has %d parameters (%s); this build implements create_G_decoder_upsampling32 with %d"
                          % (pg.size, f.net_describe("G"), ctx.count(NET_G)))
        ctx.set_params(NET_G, pg)
        bn = f.net_bn_state("G")
        if bn.size == 768:
            ctx.set_bn_state(bn)
        elif bn.size:
            raise FGError("checkpoint G carries %d BatchNorm running statistics, expected 768 (2 layers: 256 + 128 channels)"
                          % bn.size)
        else:
            # prepareNetworkForSave never strips running_mean / running_var (utils/nn_utils.lua:259-279), so a stock
            # checkpoint always has them; without them evaluate()-mode G would silently use stale statistics
            import warnings
            warnings.warn("checkpoint G holds no BatchNorm running statistics: evaluate()-mode forwards will use the "
                          "context's current ones (training-mode forwards, incl. sample.lua's, are unaffected)")
        if want_D and f.kind("D") is not None:
            pd = f.net_params("D")
            if pd.size != ctx.count(NET_D):
                raise FGError("checkpoint D has %d parameters (%s); this build implements create_D32b with %d"
                              % (pd.size, f.net_describe("D"), ctx.count(NET_D)))
            ctx.set_params(NET_D, pd)
        return int(f.number("epoch")) if f.kind("epoch") == "number" else None


def save_flat_checkpoint(ctx, path, epoch=0, extra=None):
    """Everything needed to resume: parameters, Adam moments and step counters, BN running statistics."""
    w = T7Writer(path)
    for name, net in (("G", NET_G), ("D", NET_D)):
        w.add(name, ctx.get_params(net))
        m, v, t = ctx.get_adam_state(net)
        w.add("adam_%s_m" % name, m)
        w.add("adam_%s_v" % name, v)
        w.add("adam_%s_t" % name, t)
    w.add("bn_G", ctx.get_bn_state())
    w.add("epoch", epoch)
    w.add("format", "fg_b200 flat checkpoint: getParameters()-ordered vectors of create_G_decoder_upsampling32 / create_D32b")
    for k, v in (extra or {}).items():
        w.add(k, v)
    w.close()


def load_flat_checkpoint(ctx, path):
    with T7File(path) as f:
        for name, net in (("G", NET_G), ("D", NET_D)):
            ctx.set_params(net, f.tensor(name))
            ctx.set_adam_state(net, f.tensor("adam_%s_m" % name), f.tensor("adam_%s_v" % name),
                               int(f.number("adam_%s_t" % name)))
        ctx.set_bn_state(f.tensor("bn_G"))
        return int(f.number("epoch"))
