"""Device-resident dataset + on-GPU batch assembly (dataset.lua:80-117, adversarial.lua:244-249).

CPU: the numpy restatement of image.scale has the properties the `image` rock's algorithm has (2x2 box mean for the
reference's 64 -> 32, partition of unity, identity at equal size).
GPU: fg_dataset_gather against that restatement; the device index / noise streams; the device-fed train step equals
fg_train_step on the same (gathered, drawn) inputs."""
import numpy as np
import pytest

import parity_utils as PU
from oracle import oracle_data as OD


def test_scale_restatement_properties():
    rng = np.random.default_rng(1)
    x = rng.random((2, 3, 64, 64))
    y = OD.scale(x, 32, 32)
    np.testing.assert_allclose(y, x.reshape(2, 3, 32, 2, 32, 2).mean(axis=(3, 5)), rtol=1e-6)  # dataset.lua: 64 -> 32
    np.testing.assert_array_equal(OD.scale(x, 64, 64), x)
    for s, d in ((64, 32), (96, 32), (50, 32), (45, 32), (20, 32), (16, 32), (1, 4)):
        W = OD._axis_matrix(s, d)
        np.testing.assert_allclose(W.sum(axis=1), 1.0, rtol=1e-6)  # every output is a weighted mean
        assert (W >= 0).all()
    W = OD._axis_matrix(16, 32)  # enlarging: linear interpolation, end points kept
    assert W[0, 0] == 1.0 and W[31, 15] == 1.0 and np.count_nonzero(W[5]) == 2
    W = OD._axis_matrix(48, 32)  # 1.5 source pixels per output
    np.testing.assert_allclose(W[0, :3], [2 / 3, 1 / 3, 0], atol=1e-6)
    np.testing.assert_allclose(W[1, :4], [0, 1 / 3, 2 / 3, 0], atol=1e-6)
    g = OD.load_float((rng.random((2, 3, 8, 8)) * 255).astype(np.uint8), 1)
    assert g.shape == (2, 1, 8, 8) and g.max() <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("Cs,C,Hs,Ws", [(3, 3, 64, 64), (3, 1, 64, 64), (1, 1, 64, 64), (3, 3, 96, 80), (3, 3, 50, 45),
                                        (3, 3, 32, 32), (3, 3, 20, 16), (1, 1, 1, 1)])
def test_gpu_gather_matches_image_scale(Cs, C, Hs, Ws):
    import face_generator_b200 as fg
    from face_generator_b200.dataset import DeviceDataset
    rng = np.random.default_rng(Hs * 100 + Ws + C)
    N, B = 37, 16
    imgs = rng.integers(0, 256, (N, Cs, Hs, Ws), dtype=np.uint8)
    ctx = fg.Context(0, max_batch=B, channels=C)
    ds = DeviceDataset(ctx, imgs, chunk=10)
    assert ds.size() == N
    idx = rng.integers(0, N, B)
    got = ds.gather(idx)
    ref = OD.gather(imgs, idx, C)
    assert got.shape == (B, C, 32, 32)
    assert np.abs(got - ref).max() < 2e-6  # [0,1] data, fp32 accumulation order is the only difference
    with pytest.raises(fg.FGError):
        ds.gather([N])
    ds.close()
    ctx.close()


@pytest.mark.gpu
def test_gpu_index_and_noise_streams():
    import face_generator_b200 as fg
    from face_generator_b200.dataset import DeviceDataset, noise_uniform
    ctx = fg.Context(0, max_batch=256, channels=3)
    ds = DeviceDataset(ctx, np.zeros((1000, 3, 8, 8), np.uint8))
    a, b, c = ds.draw(5, 256), ds.draw(5, 256), ds.draw(6, 256)
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, c) and a.min() >= 0 and a.max() < 1000
    many = np.concatenate([ds.draw(s, 256) for s in range(100, 140)])
    assert abs(many.mean() - 499.5) < 10 and len(np.unique(many)) > 990
    n1, n2 = noise_uniform(ctx, 9, (1 << 18,)), noise_uniform(ctx, 9, (1 << 18,))
    np.testing.assert_array_equal(n1, n2)
    assert n1.min() >= -1.0 and n1.max() < 1.0 and abs(n1.mean()) < 5e-3 and abs(n1.var() - 1 / 3) < 5e-3
    ds.close()
    ctx.close()


@pytest.mark.gpu
def test_gpu_device_fed_step_equals_host_fed_step():
    """fg_train_step_dataset == fg_train_step on (gather(draw(4s)), uniform(4s+1), uniform(4s+2)) with the same mask seed."""
    import face_generator_b200 as fg
    from face_generator_b200.dataset import DeviceDataset, noise_uniform
    from face_generator_b200.lib import NET_D, NET_G
    B, C, seed = 16, 3, 21
    rng = np.random.default_rng(4)
    imgs = rng.integers(0, 256, (200, 3, 64, 64), dtype=np.uint8)
    case = PU.make_case(B, C, seed=88, init="smooth")
    hyper = fg.hyper_default()
    res = []
    for mode in ("device", "host"):
        ctx = fg.Context(0, max_batch=B, channels=C)
        ctx.set_params(NET_G, case["PG"])
        ctx.set_params(NET_D, case["PD"])
        ds = DeviceDataset(ctx, imgs)
        if mode == "device":
            st = ds.train_step(hyper, B, seed)
        else:
            real = ds.gather(ds.draw(4 * seed, B // 2))
            nD = noise_uniform(ctx, 4 * seed + 1, (B // 2, 100))
            nG = noise_uniform(ctx, 4 * seed + 2, (B, 100))
            st = ctx.train_step(hyper, B, real, nD, nG, None, None, seed)
        res.append((st, ctx.get_grads(NET_D), ctx.get_grads(NET_G), ctx.get_params(NET_G)))
        ds.close()
        ctx.close()
    (s1, gd1, gg1, p1), (s2, gd2, gg2, p2) = res
    assert abs(s1["loss_D"] - s2["loss_D"]) < 1e-5 and abs(s1["loss_G"] - s2["loss_G"]) < 1e-5 and s1["conf"] == s2["conf"]
    assert PU.relerr(gd1, gd2) < 2e-5 and PU.relerr(gg1, gg2) < 2e-5  # split-K atomics order only
    assert PU.relerr(p1, p2) < 1e-5


def test_scale_restatement_vs_torch_interpolate():
    """Where PyTorch has the same rule: integer shrink factors = 'area', enlarging = bilinear with align_corners
    (scale (src-1)/(dst-1), image.c Main_scaleLinear_rowcol)."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(3)
    for H, W in ((64, 64), (96, 128), (32, 160)):
        x = rng.random((2, 3, H, W))
        ref = F.interpolate(torch.tensor(x), size=(32, 32), mode="area").numpy()
        np.testing.assert_allclose(OD.scale(x, 32, 32), ref, rtol=1e-6, atol=1e-7)
    for H, W in ((16, 16), (20, 9), (31, 2)):
        x = rng.random((2, 1, H, W))
        ref = F.interpolate(torch.tensor(x), size=(32, 32), mode="bilinear", align_corners=True).numpy()
        np.testing.assert_allclose(OD.scale(x, 32, 32), ref, rtol=1e-5, atol=1e-6)
