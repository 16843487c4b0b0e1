"""TEST INFRASTRUCTURE ONLY.  numpy (float64) restatement of the reference's non-default optimizers, which ARE pinned
in-repo (interruptable_optimizers.lua): interruptableAdagrad :7-46 and interruptableSgd :97-167, with the options
train.lua actually sets (OPTSTATE at train.lua:180-188: adagrad = {}, sgd = {learningRate, momentum})."""
import numpy as np


def adagrad_step(x, g, state, lr=1e-3, lrd=0.0):
    """state: dict with paramVariance (array or None), evalCounter.  In place on x."""
    nevals = state.get("evalCounter", 0)
    clr = lr / (1 + nevals * lrd)                                   # :29
    if state.get("paramVariance") is None:                          # :32-35
        state["paramVariance"] = np.zeros_like(x)
    state["paramVariance"] += g * g                                 # :36
    std = np.sqrt(state["paramVariance"]) + 1e-10                   # :37-38
    x -= clr * g / std                                              # :38
    state["evalCounter"] = nevals + 1                               # :41


def sgd_step(x, g, state, lr=1e-3, lrd=0.0, wd=0.0, mom=0.0, damp=None, nesterov=False):
    """state: dict with dfdx (momentum buffer or None), evalCounter.  In place on x."""
    damp = mom if damp is None else damp                            # :104
    nevals = state.get("evalCounter", 0)
    g = g.copy()
    if wd != 0:                                                     # :124-125
        g += wd * x
    if mom != 0:                                                    # :136-147
        if state.get("dfdx") is None:
            state["dfdx"] = g.copy()
        else:
            state["dfdx"] = state["dfdx"] * mom + (1 - damp) * g
        g = g + mom * state["dfdx"] if nesterov else state["dfdx"]
    clr = lr / (1 + nevals * lrd)                                   # :150
    x -= clr * g                                                    # :160
    state["evalCounter"] = nevals + 1                               # :164
