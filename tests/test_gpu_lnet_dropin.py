"""L-net drop-in: the UNMODIFIED control flow of train.lua:134-152 + adversarial.lua:83-288
+ interruptable_optimizers.lua:49-94, transcribed line by line, running on the fused modules of
face_generator_b200/nn.py (the executable mirror of lua/b200.lua).  torch CUDA tensors stand in for cutorch's
CudaTensors (device memory + the elementwise ops the Lua host applies to PARAMETERS / GRAD_PARAMETERS); every
model FLOP goes through libfg_b200.so.  Checked against the fused fg_train_step on the same inputs."""
import numpy as np
import pytest

import parity_utils as PU

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def stock_bce_forward(x, t):  # nn.BCECriterion on the CPU (train.lua:148), 2015 nn: eps = 1e-12
    x, t = np.asarray(x, np.float32).ravel(), np.asarray(t, np.float32).ravel()
    return float(-np.mean(t * np.log(x + 1e-12) + (1 - t) * np.log(1 - x + 1e-12)))


def stock_bce_backward(x, t):
    x, t = np.asarray(x, np.float32), np.asarray(t, np.float32).reshape(np.asarray(x).shape)
    return (-(t - x) / ((1 - x + 1e-12) * x + 1e-12) / x.size).astype(np.float32)


def stock_interruptableAdam(opfunc, x, config, state=None):
    """interruptable_optimizers.lua:49-94, tensor op by tensor op (torch stands in for cutorch)"""
    state = config if state is None else state                                   # :52
    lr = config.get("learningRate", 0.001)                                       # :53
    beta1, beta2, epsilon = config.get("beta1", 0.9), config.get("beta2", 0.999), config.get("epsilon", 1e-8)
    fx, dfdx = opfunc(x)                                                          # :60
    if fx is False:                                                               # :64-66
        return False
    state["t"] = state.get("t", 0)
    if "m" not in state:
        state["m"], state["v"], state["denom"] = torch.zeros_like(dfdx), torch.zeros_like(dfdx), torch.zeros_like(dfdx)
    state["t"] += 1                                                               # :78
    state["m"].mul_(beta1).add_(dfdx, alpha=1 - beta1)                            # :81
    state["v"].mul_(beta2).addcmul_(dfdx, dfdx, value=1 - beta2)                  # :82
    state["denom"].copy_(state["v"]).sqrt_().add_(epsilon)                        # :84
    biasCorrection1, biasCorrection2 = 1 - beta1 ** state["t"], 1 - beta2 ** state["t"]
    stepSize = lr * np.sqrt(biasCorrection2) / biasCorrection1                     # :88 (Lua double)
    x.addcdiv_(state["m"], state["denom"], value=-stepSize)                       # :90
    return x, [fx]


def run_reference_flow(fg, case, B, C, optimizer, use_stock_bce):
    from face_generator_b200 import nn as NN
    from face_generator_b200.lib import NET_D, NET_G
    OPT = dict(D_L1=0.0, D_L2=1e-4, G_L1=0.0, G_L2=0.0, D_clamp=1.0, G_clamp=5.0)
    ctx = fg.Context(0, max_batch=B, channels=C)
    # ---- train.lua:134-138: MODELS.create_D / create_G (swapped factories) + NN_UTILS.initializeWeights --------
    MODEL_D, MODEL_G = NN.FusedD(ctx), NN.FusedG(ctx)
    for model, flat in ((MODEL_D, case["PD"]), (MODEL_G, case["PG"])):
        o = 0
        for m in model.modules:  # nn_utils.lua:21-28 writes every top-level module's .weight / .bias
            for t in (m.weight, m.bias):
                if t is not None:
                    t.copy_(torch.from_numpy(flat[o:o + t.numel()].reshape(tuple(t.shape))))
                    o += t.numel()
        assert o == flat.size
    # ---- train.lua:142-145 activateCuda (clone + :cuda() are no-ops on a fused net), :148, :151-152 ------------
    MODEL_D, MODEL_G = MODEL_D.clone().cuda(), MODEL_G.clone().cuda()
    crit = NN.BCECriterion(ctx)
    cf = (lambda o, t: stock_bce_forward(o, t)) if use_stock_bce else crit.forward
    cb = (lambda o, t: stock_bce_backward(o, t)) if use_stock_bce else crit.backward
    PARAMETERS_D, GRAD_PARAMETERS_D = MODEL_D.getParameters()
    PARAMETERS_G, GRAD_PARAMETERS_G = MODEL_G.getParameters()
    NN.register_parameters(MODEL_D, PARAMETERS_D)
    NN.register_parameters(MODEL_G, PARAMETERS_G)
    # the flat tensors are NEW storages (Module.flatten): not the library's own buffers any more
    assert PARAMETERS_D.data_ptr() != ctx.lib.fg_params_ptr(ctx.h, NET_D)
    OPTSTATE = {"D": {}, "G": {}}
    inputs = np.concatenate([case["real"], np.zeros_like(case["real"])])
    targets = np.zeros(B, np.float32)
    log = {}

    def fevalD(x):                                                                # adversarial.lua:83-179
        assert x is PARAMETERS_D
        GRAD_PARAMETERS_D.zero_()                                                 # :92
        outputs = MODEL_D.forward(inputs)                                         # :95
        f = cf(outputs, targets)                                                  # :96
        df_do = cb(outputs, targets)                                              # :99
        MODEL_D.backward(inputs, df_do)                                           # :100
        if OPT["D_L1"] != 0 or OPT["D_L2"] != 0:                                  # :103-109
            f = f + OPT["D_L1"] * float(torch.norm(PARAMETERS_D, 1))
            f = f + OPT["D_L2"] * float(torch.norm(PARAMETERS_D, 2)) ** 2 / 2
            GRAD_PARAMETERS_D.add_(torch.sign(PARAMETERS_D).mul_(OPT["D_L1"]) + PARAMETERS_D.clone().mul_(OPT["D_L2"]))
        if OPT["D_clamp"] != 0:                                                   # :121-123
            GRAD_PARAMETERS_D.clamp_(-OPT["D_clamp"], OPT["D_clamp"])
        log["f_D"], log["outputs_D"] = f, outputs.copy()
        return f, GRAD_PARAMETERS_D                                               # :171 (gate open: D_maxAcc = 1.01)

    def fevalG_on_D(x):                                                           # adversarial.lua:187-231
        assert x is PARAMETERS_G
        GRAD_PARAMETERS_G.zero_()                                                 # :193
        samples = MODEL_G.forward(noiseInputs)                                    # :202
        outputs = MODEL_D.forward(samples)                                        # :204
        f = cf(outputs, targets)                                                  # :205
        df_samples = cb(outputs, targets)                                         # :208
        df_do = MODEL_D.backward(samples, df_samples)                             # :209-210 (modules[1].gradInput)
        MODEL_G.backward(noiseInputs, df_do)                                      # :214
        if OPT["G_L1"] != 0 or OPT["G_L2"] != 0:                                  # :218-224 (incl. the :223 quirk)
            f = f + OPT["G_L1"] * float(torch.norm(PARAMETERS_G, 1))
            f = f + OPT["G_L2"] * float(torch.norm(PARAMETERS_G, 2)) ** 2 / 2
            GRAD_PARAMETERS_G.add_(torch.sign(PARAMETERS_G).mul_(OPT["G_L2"]) + PARAMETERS_G.clone().mul_(OPT["G_L2"]))
        if OPT["G_clamp"] != 0:                                                   # :226-228
            GRAD_PARAMETERS_G.clamp_(-OPT["G_clamp"], OPT["G_clamp"])
        log["f_G"] = f
        return f, GRAD_PARAMETERS_G

    # ---- D step (adversarial.lua:240-268) ----
    Bh = B // 2
    targets[:Bh] = 1                                                              # Y_NOT_GENERATOR
    MODEL_D.masks = None
    samples = MODEL_G.forward(case["noise_D"])                                    # :252 createImages (train mode)
    inputs[Bh:] = samples
    targets[Bh:] = 0                                                              # Y_GENERATOR
    MODEL_D.masks = case["masks_D"]
    optimizer(fevalD, PARAMETERS_D, OPTSTATE["D"])                                # :264
    gD = GRAD_PARAMETERS_D.cpu().numpy().copy()
    # ---- G step (adversarial.lua:275-288) ----
    noiseInputs = case["noise_G"]                                                 # :276
    targets[:] = 1                                                                # :277
    MODEL_D.masks = case["masks_G"]
    optimizer(fevalG_on_D, PARAMETERS_G, OPTSTATE["G"])                           # :284
    res = dict(PD=PARAMETERS_D.cpu().numpy(), PG=PARAMETERS_G.cpu().numpy(), gD=gD, gG=GRAD_PARAMETERS_G.cpu().numpy(),
               f_D=log["f_D"], f_G=log["f_G"], tD=OPTSTATE["D"]["t"], tG=OPTSTATE["G"]["t"])
    # a gated step: opfunc returns false, false (adversarial.lua:177) -> the optimizer returns false, x / t untouched
    before, t0 = PARAMETERS_D.clone(), OPTSTATE["D"]["t"]
    assert optimizer(lambda x: (False, False), PARAMETERS_D, OPTSTATE["D"]) is False
    assert torch.equal(before, PARAMETERS_D) and OPTSTATE["D"]["t"] == t0
    ctx.close()
    return res


@pytest.mark.parametrize("which", ["stock", "b200"])
def test_unmodified_adversarial_flow_equals_fused_step(which):
    """stock: stock interruptableAdam + stock CPU BCECriterion on aliased flat tensors (nothing but the two model
    factories is swapped); b200: b200.interruptableAdam + b200.BCECriterion drop-ins.  Both must give the fused
    fg_train_step's losses, gradients and parameters."""
    import face_generator_b200 as fg
    from face_generator_b200 import nn as NN
    from face_generator_b200.lib import NET_D, NET_G
    B, C = 8, 3
    case = PU.make_case(B, C, seed=61, init="smooth")  # no PReLU kinks: the paths may only differ by summation order
    opt = stock_interruptableAdam if which == "stock" else NN.interruptableAdam
    got = run_reference_flow(fg, case, B, C, opt, use_stock_bce=(which == "stock"))
    ctx = fg.Context(0, max_batch=B, channels=C)
    ctx.set_params(NET_G, case["PG"])
    ctx.set_params(NET_D, case["PD"])
    st = ctx.train_step(fg.hyper_default(), B, case["real"], case["noise_D"], case["noise_G"], case["masks_D"], case["masks_G"])
    ref = dict(PD=ctx.get_params(NET_D), PG=ctx.get_params(NET_G), gD=ctx.get_grads(NET_D), gG=ctx.get_grads(NET_G))
    ctx.close()
    assert got["tD"] == 1 and got["tG"] == 1
    assert abs(got["f_D"] - st["loss_D"]) < 1e-5 * max(1, abs(st["loss_D"]))
    assert abs(got["f_G"] - st["loss_G"]) < 1e-4 * max(1, abs(st["loss_G"]))
    assert PU.relerr(got["gD"], ref["gD"]) < 2e-5 and PU.relerr(got["gG"], ref["gG"]) < 1e-4
    for k in ("PD", "PG"):  # sign flips of noise-level gradients move a parameter by 2*lr
        assert np.abs(got[k] - ref[k]).max() < 2.1e-3
        assert np.mean(np.abs(got[k] - ref[k]) > 1e-5) < 1e-3
