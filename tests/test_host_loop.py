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
