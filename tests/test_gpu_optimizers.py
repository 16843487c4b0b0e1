"""GPU: --D_optmethod / --G_optmethod adagrad | sgd (train.lua:38-39; interruptable_optimizers.lua:7-46, :97-167)
through fg_optim_step, against the numpy restatement oracle/oracle_optim.py.  The gradients come from a real D / G
backward; after the call fg_get_grads returns the post-penalty, post-clamp gradient the optimizer consumed."""
import numpy as np
import pytest

import parity_utils as PU
from oracle import oracle_optim as OO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method,mom", [("adagrad", 0.0), ("sgd", 0.0), ("sgd", 0.9), ("adam", 0.0)])
def test_optim_step_rules(method, mom):
    import face_generator_b200 as fg
    from face_generator_b200.lib import NET_D, NET_G
    B, C = 8, 3
    case = PU.make_case(B, C, seed=91)
    rng = np.random.default_rng(3)
    ctx = fg.Context(0, max_batch=B, channels=C)
    lr = {"adagrad": 1e-3, "sgd": 0.02, "adam": 1e-3}[method]  # OPTSTATE defaults (train.lua:180-191, :21-24)
    hyper = fg.hyper_default(lr_D=lr, lr_G=lr)
    for net, P in ((NET_D, case["PD"]), (NET_G, case["PG"])):
        ctx.set_params(net, P)
        ctx.set_optimizer(net, method, mom)
        x = P.astype(np.float64)
        state = {}
        m = np.zeros_like(x)
        v = np.zeros_like(x)
        for it in range(3):
            ctx.zero_grads(net)
            if net == NET_D:
                ctx.D_forward(rng.random((B, C, 32, 32)).astype(np.float32), masks=case["masks_D"])
                ctx.D_backward(rng.standard_normal(B).astype(np.float32), want_dimages=False)
            else:
                ctx.G_forward(case["noise_G"], want_images=False)
                ctx.G_backward(rng.standard_normal((B, C, 32, 32)).astype(np.float32))
            ctx.optim_step(net, hyper)
            g = ctx.get_grads(net).astype(np.float64)  # penalty (D_L2 = 1e-4) + clamp already applied
            assert np.abs(g).max() <= (1.0 if net == NET_D else 5.0) + 1e-6
            if method == "adagrad":
                OO.adagrad_step(x, g, state, lr=lr)
            elif method == "sgd":
                OO.sgd_step(x, g, state, lr=lr, mom=mom)
            else:  # interruptableAdam :69-90, for reference
                t = it + 1
                m = 0.9 * m + 0.1 * g
                v = 0.999 * v + 0.001 * g * g
                x -= lr * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
            got = ctx.get_params(net)
            assert np.abs(got - x).max() < 3e-6 * max(1.0, lr / 1e-3), (method, net, it, np.abs(got - x).max())
            _, _, t_dev = ctx.get_adam_state(net)
            assert t_dev == it + 1  # state.evalCounter / state.t
    ctx.close()


def test_train_step_with_sgd_and_adagrad_runs_and_differs_from_adam():
    import face_generator_b200 as fg
    from face_generator_b200.lib import NET_D, NET_G
    B, C = 8, 3
    case = PU.make_case(B, C, seed=92)
    out = {}
    for method in ("adam", "adagrad", "sgd"):
        ctx = fg.Context(0, max_batch=B, channels=C)
        ctx.set_params(NET_G, case["PG"])
        ctx.set_params(NET_D, case["PD"])
        ctx.set_optimizer(NET_D, method, 0.5)
        ctx.set_optimizer(NET_G, method, 0.5)
        lr = 0.02 if method == "sgd" else 1e-3
        st = ctx.train_step(fg.hyper_default(lr_D=lr, lr_G=lr), B, case["real"], case["noise_D"], case["noise_G"],
                            case["masks_D"], case["masks_G"])
        assert np.isfinite(st["loss_D"]) and np.isfinite(st["loss_G"]) and st["t_D"] == 1 and st["t_G"] == 1
        out[method] = (ctx.get_params(NET_D), ctx.get_grads(NET_D))
        ctx.close()
    # same D gradient (the D step happens before any update), different update rule
    assert PU.relerr(out["sgd"][1], out["adam"][1]) < 1e-4
    g = out["adam"][1].astype(np.float64)
    np.testing.assert_allclose(out["sgd"][0], case["PD"].astype(np.float64) - 0.02 * g, atol=3e-6)
    big = np.abs(g) > 1e-4 * np.abs(g).max()
    np.testing.assert_allclose(out["adagrad"][0][big], (case["PD"].astype(np.float64) - 1e-3 * np.sign(g))[big], atol=3e-6)
