"""CPU: the numpy restatement of interruptableAdagrad / interruptableSgd (oracle/oracle_optim.py) against
torch.optim.Adagrad / torch.optim.SGD, which descend from the same `optim` package the reference copied
(interruptable_optimizers.lua:1-4)."""
import numpy as np
import pytest
import torch

from oracle import oracle_optim as OO


def _run_torch(opt_ctor, x0, grads):
    p = torch.tensor(x0.copy(), requires_grad=True)
    opt = opt_ctor([p])
    for g in grads:
        p.grad = torch.tensor(g)
        opt.step()
    return p.detach().numpy()


def test_adagrad_matches_torch():
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal(500)
    grads = [rng.standard_normal(500) for _ in range(5)]
    x, st = x0.copy(), {}
    for g in grads:
        OO.adagrad_step(x, g, st, lr=1e-3)  # OPTSTATE.adagrad = {} -> lr 1e-3, no decay (train.lua:181)
    ref = _run_torch(lambda ps: torch.optim.Adagrad(ps, lr=1e-3, lr_decay=0, eps=1e-10), x0, grads)
    np.testing.assert_allclose(x, ref, rtol=1e-12, atol=1e-15)
    assert st["evalCounter"] == 5


@pytest.mark.parametrize("mom", [0.0, 0.5, 0.9])
def test_sgd_matches_torch(mom):
    rng = np.random.default_rng(2)
    x0 = rng.standard_normal(500)
    grads = [rng.standard_normal(500) for _ in range(5)]
    x, st = x0.copy(), {}
    for g in grads:
        OO.sgd_step(x, g, st, lr=0.02, mom=mom)  # --D_SGD_lr 0.02, dampening defaults to the momentum (:104-105)
    ref = _run_torch(lambda ps: torch.optim.SGD(ps, lr=0.02, momentum=mom, dampening=mom), x0, grads)
    np.testing.assert_allclose(x, ref, rtol=1e-12, atol=1e-15)
