"""Host-side mirror of the reference's nn.Module surface for the hot path (same names, argument
meaning and error behaviour as the Lua shims in face_generator_b200/lua/b200.lua).

  b200.FusedG  <- MODELS.create_G(...)            models.lua:87-93  (create_G_decoder_upsampling32 :57-81)
  b200.FusedD  <- MODELS.create_D(...)            models.lua:98-104 (create_D32b :382-416)
  BCECriterion <- nn.BCECriterion()               train.lua:148
  interruptableAdam                               interruptable_optimizers.lua:49-94
"""
import numpy as np

from .lib import NET_D, NET_G, Context, FGError


class Module:
    def __init__(self, ctx: Context, net: int):
        self.ctx, self.net = ctx, net
        self.train = True
        self.output = None
        self.gradInput = None

    # nn.Module protocol ---------------------------------------------------------------------
    def training(self):
        self.train = True
        return self

    def evaluate(self):
        self.train = False
        return self

    def getParameters(self):
        """(flat params, flat gradParams) copies; the device buffers stay authoritative
        (Lua aliases them as CudaTensors through fg_params_ptr/fg_grads_ptr)."""
        return self.ctx.get_params(self.net), self.ctx.get_grads(self.net)

    def setParameters(self, flat):
        self.ctx.set_params(self.net, flat)

    def zeroGradParameters(self):
        self.ctx.zero_grads(self.net)


class FusedG(Module):
    def __init__(self, ctx):
        super().__init__(ctx, NET_G)

    def forward(self, noise):
        self.output = self.ctx.G_forward(noise, training=self.train)
        return self.output

    def backward(self, noise, gradOutput, want_gradInput=False):
        if self.output is None or len(noise) != len(self.output):
            raise FGError("FusedG.backward: call forward on the same input first")
        self.gradInput = self.ctx.G_backward(gradOutput, want_dnoise=want_gradInput)
        return self.gradInput


class FusedD(Module):
    def __init__(self, ctx):
        super().__init__(ctx, NET_D)
        self.masks = None  # parity mode: explicit dropout keep flags [B][1984]; None => in-kernel RNG
        self.seed = 0

    def forward(self, images):
        self.output = self.ctx.D_forward(images, masks=self.masks, training=self.train, seed=self.seed).reshape(-1, 1)
        return self.output

    def backward(self, images, gradOutput, want_wgrad=True):
        if self.output is None or len(images) != len(self.output):
            raise FGError("FusedD.backward: call forward on the same input first")
        self.gradInput = self.ctx.D_backward(np.asarray(gradOutput).reshape(-1), want_wgrad=want_wgrad)
        return self.gradInput


class BCECriterion:
    def __init__(self, ctx):
        self.ctx = ctx

    def forward(self, x, t):
        return self.ctx.bce_forward(x, t)

    def backward(self, x, t):
        return self.ctx.bce_backward(x, t).reshape(np.asarray(x).shape)


def interruptableAdam(opfunc, module: Module, hyper, grad_scale=1.0):
    """opfunc() -> f or False.  Like interruptable_optimizers.lua:64-66 the step is skipped
    (and t not advanced) when opfunc returns False; otherwise penalty/clamp were already applied
    by the caller in the reference -- here they are fused into the step (fg_optim_step)."""
    fx = opfunc()
    if fx is False:
        return False
    module.ctx.optim_step(module.net, hyper, grad_scale)
    return fx
