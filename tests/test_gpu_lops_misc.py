"""GPU: the resampling / pooling / dropout / sigmoid L-ops (NCHW at the nn.Module boundary) against the oracle.
These are exact data movements or single fp32 operations, so the bar is bit-exact / 1e-6."""
import numpy as np
import pytest

import parity_utils as PU
from oracle import oracle as O
from oracle import oracle_c2f as OC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import face_generator_b200 as fg
    c = fg.Context(0, max_batch=8, channels=3)
    yield c
    c.close()


@pytest.mark.parametrize("shape", [(3, 5, 8, 8), (2, 128, 16, 16), (1, 1, 2, 2)])
def test_upsample2(ctx, shape):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape).astype(np.float32)
    y = ctx.upsample2_forward(x)
    np.testing.assert_array_equal(y, O.f32.up2_fwd(x))  # SpatialUpSamplingNearest(2), models.lua:63
    dy = rng.standard_normal(y.shape).astype(np.float32)
    assert PU.relerr(ctx.upsample2_backward(dy), O.f64.up2_bwd(dy)) < 1e-6


@pytest.mark.parametrize("shape", [(3, 5, 8, 8), (2, 64, 32, 32), (1, 2, 2, 2)])
def test_avgpool2(ctx, shape):
    rng = np.random.default_rng(2)
    x = rng.standard_normal(shape).astype(np.float32)
    y = ctx.avgpool2_forward(x)
    assert PU.relerr(y, O.f64.avgpool2_fwd(x)) < 1e-6  # SpatialAveragePooling(2,2,2,2), models.lua:388
    dy = rng.standard_normal(y.shape).astype(np.float32)
    np.testing.assert_array_equal(ctx.avgpool2_backward(dy), O.f32.avgpool2_bwd(dy))


@pytest.mark.parametrize("shape", [(3, 5, 8, 8), (2, 64, 32, 32), (1, 2, 2, 2)])
def test_maxpool2(ctx, shape):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(shape).astype(np.float32)
    x[0, 0, 0, :2] = 7.0  # a tie inside the first window: the first element in row-major order must win
    ref_y, arg = OC.f32.maxpool2_fwd(x)
    np.testing.assert_array_equal(ctx.maxpool2_forward(x), ref_y)  # SpatialMaxPooling(2,2), models_c2f.lua:251
    dy = rng.standard_normal(ref_y.shape).astype(np.float32)
    np.testing.assert_array_equal(ctx.maxpool2_backward(x, dy), OC.f32.maxpool2_bwd(dy, arg))


def test_dropout_both_kinds(ctx):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((4, 16, 8, 8)).astype(np.float32)
    m = (rng.random(x.shape) < 0.5).astype(np.float32)
    # nn.Dropout v2: mask / (1-p) in training, identity in evaluate()        (models.lua:408, models_c2f.lua:258)
    np.testing.assert_array_equal(ctx.dropout_forward(x, m, 0.5), x * m * 2.0)
    np.testing.assert_array_equal(ctx.dropout_backward(x, m, 0.5), x * m * 2.0)
    np.testing.assert_array_equal(ctx.dropout_forward(x, None, 0.5), x)
    # nn.SpatialDropout: one flag per (n, c) plane, no rescale in training, (1-p) in evaluate()  (models.lua:387)
    ms = (rng.random((4, 16)) < 0.8).astype(np.float32)
    np.testing.assert_array_equal(ctx.dropout_forward(x, ms, 0.2, spatial=True), x * ms[:, :, None, None])
    assert PU.relerr(ctx.dropout_forward(x, None, 0.2, spatial=True), x * np.float32(0.8)) < 1e-7
    # 2-D input (after nn.View): [N][F]
    x2 = rng.standard_normal((4, 512)).astype(np.float32)
    m2 = (rng.random(x2.shape) < 0.5).astype(np.float32)
    np.testing.assert_array_equal(ctx.dropout_forward(x2, m2, 0.5), x2 * m2 * 2.0)


def test_dropout_mask_statistics(ctx):
    for p in (0.2, 0.5):
        m = ctx.dropout_mask(1 << 20, p, seed=5)
        assert set(np.unique(m)) <= {0.0, 1.0}
        assert abs(m.mean() - (1 - p)) < 3e-3
    assert not np.array_equal(ctx.dropout_mask(4096, 0.5, seed=1), ctx.dropout_mask(4096, 0.5, seed=2))
    np.testing.assert_array_equal(ctx.dropout_mask(4096, 0.5, seed=1), ctx.dropout_mask(4096, 0.5, seed=1))


def test_sigmoid(ctx):
    rng = np.random.default_rng(6)
    x = (rng.standard_normal(5000) * 6).astype(np.float32)
    y = ctx.sigmoid_forward(x)
    assert PU.relerr(y, 1.0 / (1.0 + np.exp(-x.astype(np.float64)))) < 1e-6
    dy = rng.standard_normal(5000).astype(np.float32)
    assert PU.relerr(ctx.sigmoid_backward(y, dy), dy.astype(np.float64) * y * (1.0 - y.astype(np.float64))) < 1e-6
