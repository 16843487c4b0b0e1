"""Host-side mirror of adversarial.lua's loop body ("trainBatch", adversarial.lua:54-300).

train_batch()        the fused L-step call (fg_train_step): what adversarial_b200.lua uses.
train_batch_modules() the same iteration composed from the L-net calls exactly like the reference's
                     fevalD / fevalG_on_D closures -- used to show the two levels agree.
"""
import numpy as np

from .lib import Context
from .nn import BCECriterion, FusedD, FusedG, interruptableAdam


def create_noise_inputs(n, rng, noise_dim=100):
    """NN_UTILS.createNoiseInputs (utils/nn_utils.lua:35-39): U(-1,1)."""
    return rng.uniform(-1.0, 1.0, (n, noise_dim)).astype(np.float32)


def train_batch(ctx: Context, hyper, real, noise_D, noise_G, masks_D=None, masks_G=None, seed=0, want_stats=True):
    B = 2 * (real.shape[0] if hasattr(real, "shape") else 0) or None
    assert B is not None
    return ctx.train_step(hyper, B, real, noise_D, noise_G, masks_D, masks_G, seed, want_stats)


def train_batch_modules(ctx: Context, hyper, real, noise_D, noise_G, masks_D, masks_G):
    G, D, crit = FusedG(ctx), FusedD(ctx), BCECriterion(ctx)
    Bh = real.shape[0]
    B = 2 * Bh
    out = {}
    # ---- D step (adversarial.lua:240-268) ----
    samples = G.forward(noise_D)  # createImages: G in training mode
    inputs = np.concatenate([real, samples]).astype(np.float32)
    targets = np.concatenate([np.ones(Bh), np.zeros(Bh)]).astype(np.float32)

    def fevalD():
        D.zeroGradParameters()
        D.masks = masks_D
        outputs = D.forward(inputs)
        f = crit.forward(outputs, targets)
        D.backward(inputs, crit.backward(outputs, targets), want_wgrad=True)
        out["loss_D_bce"] = f
        out["outputs_D"] = outputs.copy()
        return f

    interruptableAdam(fevalD, D, hyper)
    out["grad_D"] = ctx.get_grads(D.net)
    # ---- G step (adversarial.lua:275-288) ----
    targets1 = np.ones(B, np.float32)

    def fevalG_on_D():
        G.zeroGradParameters()
        samples = G.forward(noise_G)
        D.masks = masks_G
        outputs = D.forward(samples)
        f = crit.forward(outputs, targets1)
        df_do = D.backward(samples, crit.backward(outputs, targets1), want_wgrad=False)
        G.backward(noise_G, df_do)
        out["loss_G"] = f
        return f

    interruptableAdam(fevalG_on_D, G, hyper)
    out["grad_G"] = ctx.get_grads(G.net)
    return out
