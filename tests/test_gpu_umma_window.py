"""Hardware semantics the haloed-tile convolution kernel relies on: a tcgen05 A operand may be a SHIFTED window of a
larger 128B-swizzled shared-memory tile when (i) the 8-row core groups stay a multiple of 1024 B apart and (ii) the
swizzle XOR is derived from the absolute shared-memory address -- which is what the hardware does: the window
works with the descriptor's base-offset field left at 0 (measured: setting it to (addr >> 7) & 7 is WRONG for a
window that starts at a non-1024-aligned row), for any row pitch of the surrounding tile."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_shifted_window_operand():
    import face_generator_b200 as fg
    from face_generator_b200.lib import _ptr
    ctx = fg.Context(0, max_batch=8, channels=3)
    x = ((np.arange(288 * 32) % 1024).astype(np.float32)).reshape(288, 32)  # TF32-exact
    dx_, di, do = ctx.dev_array(x), ctx.dev_array(np.eye(32, dtype=np.float32)), ctx.dev_array(np.zeros((128, 32), np.float32))
    out = np.empty((128, 32), np.float32)
    res = {}
    for pitch in (16, 10):
        for dy in range(3):
            for dx in range(3):
                assert ctx.lib.fg_debug_umma_window(ctx.h, dx_, di, dy, dx, pitch << 8, do) == 0, ctx.lib.fg_last_error()
                ctx.lib.fg_memcpy(ctx.h, _ptr(out), do, out.nbytes)
                rows = np.array([(m // 8 + dy) * pitch + (m % 8) + dx for m in range(128)])
                res[(pitch, dy, dx)] = bool(np.array_equal(out, x[rows]))
    ctx.close()
    assert all(res.values()), res
