"""CPU: host logic of the epoch loop (adversarial.lua:29-76): which batches an epoch consists of."""
import numpy as np
import pytest

from face_generator_b200.adversarial import epoch_batches


def lua_schedule(n_epoch, batch_size):
    """Literal transcription of `for t = 1,N_epoch,dataBatchSize do ... thisBatchSize = math.min(OPT.batchSize,
    N_epoch - t + 1) ... if thisBatchSize < 4 then break end` (adversarial.lua:54-76)."""
    out = []
    data_batch = batch_size // 2
    t = 1
    while t <= n_epoch:  # Lua numeric for: t = 1, 1+step, ... while t <= limit
        this = min(batch_size, n_epoch - t + 1)
        if this < 4:
            break
        out.append((t, this))
        t += data_batch
    return out


@pytest.mark.parametrize("n_epoch,batch", [(1000, 32), (1000, 16), (64, 32), (70, 32), (33, 16), (7, 4), (3, 4), (4, 4),
                                           (1001, 256), (20000, 256), (15, 6)])
def test_epoch_schedule_matches_reference_loop(n_epoch, batch):
    got, ref = epoch_batches(n_epoch, batch), lua_schedule(n_epoch, batch)
    assert [t for t, _ in got] == [t for t, _ in ref]
    for (_, b), (_, rb) in zip(got, ref):
        assert b == rb - rb % 2 and 4 <= b <= batch  # identical, except that odd tails are made even
    # every iteration consumes batch/2 real examples: full batches until the tail (README.md:133 vs train.lua:17)
    full = [b for _, b in got if b == batch]
    assert len(full) >= max(0, (n_epoch - batch) // (batch // 2))


def test_default_epoch_shape():
    """train.lua defaults: --batchSize 32, --N_epoch 1000 -> 63 iterations, the last two on shrinking batches."""
    s = epoch_batches(1000, 32)
    assert len(s) == 63 and s[0] == (1, 32) and s[-3][1] == 32 and s[-2] == (977, 24) and s[-1] == (993, 8)
    assert sum(b // 2 for _, b in s) == 16 * 61 + 12 + 4


class _RecordingCtx:
    """stands in for face_generator_b200.Context: records what train() feeds each fused step"""

    def __init__(self):
        self.calls = []

    def train_step(self, hyper, B, real, noise_D, noise_G, masks_D, masks_G, seed):
        self.calls.append((B, real.copy(), noise_D.copy(), noise_G.copy(), seed))
        return dict(conf=[B // 2, 0, 0, B // 2], trained_D=1)


def test_epochs_and_ranks_do_not_replay_the_same_draws():
    """adversarial.lua:245,276 draws fresh math.random indices / uniform noise on every call: two consecutive epochs
    (and two data-parallel ranks) must see different real-image indices, noise and step seeds with the defaults."""
    from face_generator_b200 import adversarial as A
    data = np.arange(64, dtype=np.float32).reshape(64, 1, 1, 1) * np.ones((1, 1, 32, 32), np.float32)
    runs = {}
    for epoch, rank in ((1, 0), (2, 0), (1, 1), (1, 0)):
        ctx = _RecordingCtx()
        acc, conf, trained = A.train(ctx, data, hyper=None, batch_size=16, n_epoch=64, epoch=epoch, rank=rank)
        assert acc == 1.0 and trained == len(ctx.calls) == len(A.epoch_batches(64, 16))
        runs.setdefault((epoch, rank), []).append(ctx.calls)
    a, b, c = runs[(1, 0)][0], runs[(2, 0)][0], runs[(1, 1)][0]
    for other in (b, c):
        assert [x[4] for x in a] != [x[4] for x in other]                      # step seeds (device RNG streams)
        assert not set(x[4] for x in a) & set(x[4] for x in other)
        assert any(not np.array_equal(x[1], y[1]) for x, y in zip(a, other))    # real-image indices
        assert all(not np.array_equal(x[2], y[2]) for x, y in zip(a, other))    # noise
    # same (epoch, rank) -> reproducible
    for x, y in zip(runs[(1, 0)][0], runs[(1, 0)][1]):
        assert x[4] == y[4] and np.array_equal(x[1], y[1]) and np.array_equal(x[3], y[3])
    # a caller-owned generator is consumed, never reseeded
    rng = np.random.default_rng(5)
    c1, c2 = _RecordingCtx(), _RecordingCtx()
    A.train(c1, data, None, 16, 64, rng=rng, epoch=1)
    A.train(c2, data, None, 16, 64, rng=rng, epoch=2)
    assert all(not np.array_equal(x[2], y[2]) for x, y in zip(c1.calls, c2.calls))
