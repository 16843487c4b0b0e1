"""GPU: scoring helpers against numpy / the oracle: sortImagesByPrediction's D scores
(utils/nn_utils.lua:90-118), findClosestNeighboursOf (sample.lua:141-159), approxParzen (adversarial_c2f.lua:305-325)."""
import numpy as np
import pytest

import c2f_utils as CU
import parity_utils as PU
from oracle import oracle as O
from oracle import oracle_c2f as OC
from oracle import oracle_data as OD

pytestmark = pytest.mark.gpu


def test_d_score_and_sort():
    import face_generator_b200 as fg
    from face_generator_b200 import scoring as S
    from face_generator_b200.lib import NET_D
    C, N, chunk = 3, 40, 16  # 16 + 16 + 8: ragged last chunk
    case = PU.make_case(8, C, seed=5)
    rng = np.random.default_rng(2)
    images = rng.random((N, C, 32, 32)).astype(np.float32)
    ctx = fg.Context(0, max_batch=chunk, channels=C)
    ctx.set_params(NET_D, case["PD"])
    preds = S.d_score(ctx, images, chunk, training=False)
    d = O.f64.D()
    ref = np.concatenate([d.forward(case["PD"], images[s:s + chunk], None, training=False) for s in range(0, N, chunk)])
    assert PU.relerr(preds, ref) < 1e-4
    best, pb = S.sort_images_by_prediction(ctx, images, False, 8, chunk, training=False)
    worst, pw = S.sort_images_by_prediction(ctx, images, True, 8, chunk, training=False)
    assert (np.diff(pb) <= 0).all() and (np.diff(pw) >= 0).all() and pb[0] == preds.max() and pw[0] == preds.min()
    np.testing.assert_array_equal(best[0], images[np.argmax(preds)])
    # sample.lua's own mode: dropout live -> scores differ from evaluate() but are reproducible per seed
    p1, p2 = S.d_score(ctx, images, chunk, training=True, seed=3), S.d_score(ctx, images, chunk, training=True, seed=3)
    np.testing.assert_allclose(p1, p2, rtol=1e-5)
    assert np.abs(p1 - preds).max() > 1e-4
    ctx.close()


@pytest.mark.parametrize("Q,N,D", [(1, 7, 3072), (5, 300, 3072), (16, 1000, 1024), (3, 50, 17)])
def test_nearest_matches_numpy(Q, N, D):
    import face_generator_b200 as fg
    from face_generator_b200 import scoring as S
    rng = np.random.default_rng(Q * 1000 + N)
    cands = rng.random((N, D)).astype(np.float32)
    queries = rng.random((Q, D)).astype(np.float32)
    queries[0] = cands[N // 2]  # exact hit
    cands[N - 1] = cands[N // 2]  # duplicate: the first (lowest index) must win, like the reference's strict `<`
    ctx = fg.Context(0, max_batch=8, channels=3)
    idx, dist = S.nearest(ctx, queries, cands)
    d2 = ((queries[:, None, :].astype(np.float64) - cands[None].astype(np.float64)) ** 2).sum(-1)
    np.testing.assert_array_equal(idx, d2.argmin(1))
    assert idx[0] == N // 2 and dist[0] == 0.0
    np.testing.assert_allclose(dist, np.sqrt(d2.min(1)), rtol=1e-5, atol=1e-6)
    ctx.close()


@pytest.mark.parametrize("Cs,C", [(3, 3), (3, 1)])
def test_dataset_nearest(Cs, C):
    import face_generator_b200 as fg
    from face_generator_b200 import scoring as S
    from face_generator_b200.dataset import DeviceDataset
    rng = np.random.default_rng(9)
    N, Q = 500, 6
    imgs = rng.integers(0, 256, (N, Cs, 64, 64), dtype=np.uint8)
    ctx = fg.Context(0, max_batch=16, channels=C)
    ds = DeviceDataset(ctx, imgs)
    train32 = OD.gather(imgs, np.arange(N), C)  # what DATASET.loadImages would hold
    queries = rng.random((Q, C, 32, 32)).astype(np.float32)
    queries[2] = train32[123].astype(np.float32) + 1e-3
    res, idx = S.find_closest_neighbours(ds, queries)
    d2 = ((queries.reshape(Q, 1, -1).astype(np.float64) - train32.reshape(1, N, -1)) ** 2).sum(-1)
    np.testing.assert_array_equal(idx, d2.argmin(1))
    assert idx[2] == 123
    np.testing.assert_allclose([r[2] for r in res], np.sqrt(d2.min(1)), rtol=1e-5, atol=5e-6)  # fp32 (x-q)^2 at |x-q| = 1e-3
    assert np.abs(res[2][1] - train32[123]).max() < 2e-6
    ds.close()
    ctx.close()


def test_c2f_approx_parzen():
    import face_generator_b200 as fg
    from face_generator_b200 import scoring as S
    from face_generator_b200.lib import NET_G
    from face_generator_b200 import layouts as LY
    C, K, n = 3, 12, 3
    case = CU.make_case(16, C, seed=6)
    rng = np.random.default_rng(1)
    diff, coarse = LY.c2f_pairs(n, C, rng)
    fine = diff + coarse
    ctx = fg.Context(0, max_batch=16, channels=C)
    net = fg.C2f(ctx)
    net.set_params(NET_G, case["PG"])
    got = S.approx_parzen(net, fine, coarse, K, np.random.default_rng(77))
    g = OC.f64.G()
    r2 = np.random.default_rng(77)
    ref = []
    for i in range(n):
        noise = r2.uniform(-1, 1, (K, 1, 32, 32)).astype(np.float32)
        cond = np.repeat(coarse[i:i + 1], K, axis=0)
        neigh = g.forward(case["PG"], noise, cond) + cond
        ref.append(np.sqrt(((neigh - fine[i]) ** 2).reshape(K, -1).sum(1)).min())
    np.testing.assert_allclose(got, ref, rtol=1e-4)
    net.close()
    ctx.close()
