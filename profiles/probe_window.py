"""diagnostic dump of the UMMA shifted-window probe: where does each 16-byte chunk of the gathered operand come from?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import face_generator_b200 as fg
from face_generator_b200.lib import _ptr
ctx = fg.Context(0, max_batch=8, channels=3)
x = ((np.arange(288 * 32) % 1024).astype(np.float32)).reshape(288, 32)
dxp, di, do = ctx.dev_array(x), ctx.dev_array(np.eye(32, dtype=np.float32)), ctx.dev_array(np.zeros((128, 32), np.float32))
out = np.empty((128, 32), np.float32)
for use, pitch in ((0, 16), (0, 10), (0, 12), (0, 9)):
    for dy in (0, 1, 2):
        for dx in range(3):
            assert ctx.lib.fg_debug_umma_window(ctx.h, dxp, di, dy, dx, use | (pitch << 8), do) == 0
            ctx.lib.fg_memcpy(ctx.h, _ptr(out), do, out.nbytes)
            rows = np.array([(m // 8 + dy) * pitch + (m % 8) + dx for m in range(128)])
            ok = np.array_equal(out, x[rows])
            line = "use_bo=%d pitch=%d dy=%d dx=%d ok=%s" % (use, pitch, dy, dx, ok)
            if not ok:
                # first group of 8 rows: for each row, the source (row mod 32, chunk) of each of the 8 chunks
                desc = []
                for m in range(8):
                    src = [(int(out[m, 4 * ch]) // 32, (int(out[m, 4 * ch]) % 32) // 4) for ch in range(8)]
                    desc.append("m%d want r%d: " % (m, rows[m] % 32) + " ".join("r%dc%d" % s for s in src))
                line += "\n    " + "\n    ".join(desc)
            print(line)
ctx.close()
