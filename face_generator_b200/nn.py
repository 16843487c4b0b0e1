"""Host-side mirror of the reference's nn.Module surface for the hot path (same names, argument
meaning and error behaviour as the Lua shims in face_generator_b200/lua/b200.lua).

  b200.FusedG  <- MODELS.create_G(...)            models.lua:87-93  (create_G_decoder_upsampling32 :57-81)
  b200.FusedD  <- MODELS.create_D(...)            models.lua:98-104 (create_D32b :382-416)
  BCECriterion <- nn.BCECriterion()               train.lua:148
  interruptableAdam                               interruptable_optimizers.lua:49-94

The L-net drop-in keeps train.lua / adversarial.lua UNMODIFIED.  What that requires of
a fused module, and what this file therefore mirrors 1:1 from b200.lua:

 * `MODEL.modules[i].weight / .bias` exist per reference layer, as views into the flat device vector, because
   NN_UTILS.initializeWeights walks `model.modules` (utils/nn_utils.lua:17-29);
 * `MODEL:getParameters()` is stock nn (train.lua:151-152 calls it on the nn.Sequential that
   NN_UTILS.activateCuda wraps around the net): Module.flatten allocates ONE new flat storage, copies the
   parameters into it and re-points `module.weight` / `module.gradWeight` at it.  The fused module notices the
   re-pointing at its next call (`weight:data()` changed) and hands the new device pointers to the library with
   fg_bind_params: from then on PARAMETERS_x / GRAD_PARAMETERS_x -- the tensors adversarial.lua:92-123 zeroes,
   penalises and clamps and interruptableAdam (interruptable_optimizers.lua:78-90) updates in place -- ARE the
   buffers the kernels read and write;
 * `opt(opfunc, x, config)` takes the flat tensor x (adversarial.lua:264, :284), returns false when opfunc does.

In this mirror a "CudaTensor" is a torch.Tensor on the GPU (PyTorch only provides device memory and the
elementwise tensor ops that cutorch provides to the Lua host; none of it is on the product's kernel path).
"""
import numpy as np

from .lib import NET_D, NET_G, Context, FGError

_I = lambda s: int(np.prod(s))
# reference layer tables: (name, weight shape | None, bias shape | None) per TOP-LEVEL module, in module order, so
# that `modules[i]` lines up with models.lua and the flat offsets follow getParameters() (weight then bias)
def G_MODULES(C):  # models.lua:57-81
    return [("nn.Linear", (8192, 100), (8192,)), ("nn.View", None, None), ("nn.PReLU", (1,), None),
            ("nn.SpatialUpSamplingNearest", None, None), ("cudnn.SpatialConvolution", (256, 128, 5, 5), (256,)),
            ("nn.SpatialBatchNormalization", (256,), (256,)), ("nn.PReLU", (1,), None),
            ("nn.SpatialUpSamplingNearest", None, None), ("cudnn.SpatialConvolution", (128, 256, 5, 5), (128,)),
            ("nn.SpatialBatchNormalization", (128,), (128,)), ("nn.PReLU", (1,), None),
            ("cudnn.SpatialConvolution", (C, 128, 3, 3), (C,)), ("nn.Sigmoid", None, None)]


def D_MODULES(C):  # models.lua:382-416
    out, cin = [], C
    for cout in (64, 128, 256, 512):
        out += [("nn.SpatialConvolution", (cout, cin, 3, 3), (cout,)), ("nn.PReLU", (1,), None),
                ("nn.SpatialDropout", None, None), ("nn.SpatialAveragePooling", None, None)]
        cin = cout
    out += [("nn.View", None, None), ("nn.Linear", (512, 2048), (512,)), ("nn.PReLU", (1,), None), ("nn.Dropout", None, None),
            ("nn.Linear", (512, 512), (512,)), ("nn.PReLU", (1,), None), ("nn.Dropout", None, None),
            ("nn.Linear", (1, 512), (1,)), ("nn.Sigmoid", None, None)]
    return out


class _DevPtr:
    """lets torch alias raw device memory without copying (the mirror of torch.CudaStorage(size, ptr))"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def alias_cuda(ptr, n, device):
    import torch
    return torch.as_tensor(_DevPtr(ptr, n), device="cuda:%d" % device)


class LayerProxy:
    """what `model.modules[i]` is for the host scripts: typename + weight/bias views into the flat vector"""

    def __init__(self, typename):
        self.typename, self.weight, self.bias, self.gradWeight, self.gradBias = typename, None, None, None, None


class Module:
    def __init__(self, ctx: Context, net: int, table):
        import torch  # the tensor library of the host in this mirror (cutorch in the Lua original)
        self.ctx, self.net = ctx, net
        self.train = True
        self.output = None
        self.gradInput = None
        self.n = ctx.count(net)
        self.table = table
        # flat parameter / gradient vectors: CudaTensors over the library's own device buffers
        self.weight = alias_cuda(ctx.lib.fg_params_ptr(ctx.h, net), self.n, ctx.device)
        self.gradWeight = alias_cuda(ctx.lib.fg_grads_ptr(ctx.h, net), self.n, ctx.device)
        self._bound = (self.weight.data_ptr(), self.gradWeight.data_ptr())
        self.modules = [LayerProxy(t[0]) for t in table]
        self._views()
        self._torch = torch

    def _views(self):
        o = 0
        for m, (_, ws, bs) in zip(self.modules, self.table):
            if ws is not None:
                m.weight, m.gradWeight = self.weight[o:o + _I(ws)].view(ws), self.gradWeight[o:o + _I(ws)].view(ws)
                o += _I(ws)
            if bs is not None:
                m.bias, m.gradBias = self.weight[o:o + _I(bs)].view(bs), self.gradWeight[o:o + _I(bs)].view(bs)
                o += _I(bs)
        assert o == self.n

    def _sync(self):
        """adopt the storage nn.Module.flatten re-pointed weight / gradWeight to (see the module docstring)"""
        now = (self.weight.data_ptr(), self.gradWeight.data_ptr())
        if now != self._bound:
            rc = self.ctx.lib.fg_bind_params(self.ctx.h, self.net, now[0], now[1])
            if rc != 0:
                raise FGError("fg_bind_params failed (%d): %s" % (rc, self.ctx.lib.fg_last_error().decode()))
            self._bound = now
            self._views()
        self._torch.cuda.current_stream().synchronize()  # host-tensor ops run on torch's stream, ours on the ctx's

    # nn.Module protocol ---------------------------------------------------------------------
    def training(self):
        self.train = True
        return self

    def evaluate(self):
        self.train = False
        return self

    def parameters(self):
        return [self.weight], [self.gradWeight]

    def getParameters(self):
        """STOCK nn.Module.getParameters -> Module.flatten (what train.lua:151-152 runs): new flat storage, copy,
        re-point the module's tensors.  Nothing here knows about the library; _sync() picks the change up."""
        torch = self._torch
        params, grads = self.parameters()
        flat_p = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=params[0].device)
        flat_g = torch.zeros_like(flat_p)
        o = 0
        for p, g in zip(params, grads):
            flat_p[o:o + p.numel()].copy_(p)
            flat_g[o:o + g.numel()].copy_(g)
            o += p.numel()
        self.weight, self.gradWeight = flat_p, flat_g  # `parameters[k]:set(flatStorage, offset, size)` on our tensors
        return flat_p, flat_g

    def zeroGradParameters(self):
        self._sync()
        self.ctx.zero_grads(self.net)

    def clone(self):  # NN_UTILS.activateCuda clones the net (nn_utils.lua:352): a fused net is one per context
        return self

    def cuda(self):  # parameters already live on the device
        return self


class FusedG(Module):
    def __init__(self, ctx):
        super().__init__(ctx, NET_G, G_MODULES(ctx.C))

    def forward(self, noise):
        self._sync()
        self.output = self.ctx.G_forward(np.ascontiguousarray(noise, np.float32), training=self.train)
        return self.output

    def backward(self, noise, gradOutput, want_gradInput=False):
        if self.output is None or len(noise) != len(self.output):
            raise FGError("FusedG.backward: call forward on the same input first")
        self._sync()
        self.gradInput = self.ctx.G_backward(np.ascontiguousarray(gradOutput, np.float32), want_dnoise=want_gradInput)
        self.ctx.sync()
        return self.gradInput


class FusedD(Module):
    def __init__(self, ctx):
        super().__init__(ctx, NET_D, D_MODULES(ctx.C))
        self.masks = None  # parity mode: explicit dropout keep flags [B][1984]; None => in-kernel RNG
        self.seed = 0

    def forward(self, images):
        self._sync()
        self.output = self.ctx.D_forward(np.ascontiguousarray(images, np.float32), masks=self.masks, training=self.train,
                                         seed=self.seed).reshape(-1, 1)
        return self.output

    def backward(self, images, gradOutput, want_wgrad=True):
        if self.output is None or len(images) != len(self.output):
            raise FGError("FusedD.backward: call forward on the same input first")
        self._sync()
        self.gradInput = self.ctx.D_backward(np.asarray(gradOutput, np.float32).reshape(-1), want_wgrad=want_wgrad)
        self.ctx.sync()
        return self.gradInput


class BCECriterion:
    def __init__(self, ctx):
        self.ctx = ctx

    def forward(self, x, t):
        return self.ctx.bce_forward(x, t)

    def backward(self, x, t):
        return self.ctx.bce_backward(x, t).reshape(np.asarray(x).shape)


_owners = {}  # data pointer of a flat parameter tensor -> fused module (b200._owners in the Lua shim)


def register_parameters(module, flat):
    """after `PARAMETERS_x = MODEL_x:getParameters()`: lets the optimizer drop-in find the module of a flat tensor"""
    _owners[flat.data_ptr()] = module


def interruptableAdam(opfunc, x, config, state=None):
    """b200.interruptableAdam(opfunc, x, config[, state]) -- same signature and return values as
    interruptable_optimizers.lua:49-94: x is the FLAT PARAMETER TENSOR (adversarial.lua:264, :284 pass
    PARAMETERS_D / PARAMETERS_G), opfunc(x) -> f, dfdx or false, false.  Returns False without touching x, the
    moments or the step counter when opfunc returns false (:64-66).  The update itself is ONE fused kernel on the
    raw device pointers (fg_adam_step: 7 streams instead of the reference's 8 cutorch kernels); the gradient
    arrives already penalised and clamped by the caller's opfunc, so no penalty / clamp is applied here.
    A FusedG / FusedD module is also accepted as `x` (then opfunc() takes no argument and penalty + clamp + Adam
    run fused on the library's own state: the L-step's optimizer, used by train_batch_modules)."""
    import torch
    if isinstance(x, Module):
        fx = opfunc()
        if fx is False:
            return False
        x.ctx.optim_step(x.net, config, 1.0)
        return x, [fx]
    state = config if state is None else state
    lr = state.get("learningRate", 1e-3)
    beta1, beta2, eps = state.get("beta1", 0.9), state.get("beta2", 0.999), state.get("epsilon", 1e-8)
    fx, dfdx = opfunc(x)
    if fx is False:
        return False
    state["t"] = state.get("t", 0) + 1
    if "m" not in state:
        state["m"], state["v"] = torch.zeros_like(dfdx), torch.zeros_like(dfdx)
    module = _owners.get(x.data_ptr())
    if module is None:
        raise FGError("b200.interruptableAdam: x is not a registered flat parameter tensor")
    torch.cuda.current_stream().synchronize()
    rc = module.ctx.lib.fg_adam_step(module.ctx.h, x.data_ptr(), dfdx.data_ptr(), state["m"].data_ptr(), state["v"].data_ptr(),
                                     x.numel(), lr, beta1, beta2, eps, state["t"], 0.0, 0.0, 0.0, 1.0)
    if rc != 0:
        raise FGError("fg_adam_step failed (%d): %s" % (rc, module.ctx.lib.fg_last_error().decode()))
    module.ctx.sync()
    return x, [fx]
