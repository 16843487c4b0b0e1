"""debug helper (not a test): localise differences between the SIMT and tcgen05 paths of G backward"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import face_generator_b200 as fg
from face_generator_b200.lib import NET_G
from face_generator_b200 import layouts as LY
import parity_utils as PU
B, C = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 3
case = PU.make_case(2 * B, C, seed=34)
rng = np.random.default_rng(7)
noise = case["noise_G"][:B]
dout = rng.standard_normal((B, C, 32, 32)).astype(np.float32)
res = {}
for impl in (0, 2):
    ctx = fg.Context(0, max_batch=max(8, B), channels=C)
    ctx.set_option("conv_impl", impl)
    ctx.set_params(NET_G, case["PG"])
    ctx.G_forward(noise)
    ctx.zero_grads(NET_G)
    ctx.G_backward(dout)
    res[impl] = dict(dz2=ctx.debug_tensor("G.dz2"), dz1=ctx.debug_tensor("G.dz1"), dz0=ctx.debug_tensor("G.dz0"), g=ctx.get_grads(NET_G))
    ctx.close()
for k, shp in (("dz2", (B, 32, 32, 128)), ("dz1", (B, 16, 16, 256)), ("dz0", (B, 8, 8, 128))):
    a, b = res[2][k].reshape(shp), res[0][k].reshape(shp)
    d = np.abs(a - b)
    i = np.unravel_index(d.argmax(), shp)
    print(k, "relerr", d.max() / np.abs(b).max(), "at", i, "tc", a[i], "simt", b[i])
    # error by (y, x) position
    e = d.max(axis=(0, 3)) / np.abs(b).max()
    print("   per-row max:", np.array2string(e.max(axis=1), precision=1, max_line_width=200))
    print("   per-col max:", np.array2string(e.max(axis=0), precision=1, max_line_width=200))
    print("   per-image max:", np.array2string(d.max(axis=(1, 2, 3)) / np.abs(b).max(), precision=1))
lay, _ = LY.G_layout(C)
for k, (o, s) in lay.items():
    n = int(np.prod(s))
    a, b = res[2]["g"][o:o + n], res[0]["g"][o:o + n]
    print("grad", k, "relerr", np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
