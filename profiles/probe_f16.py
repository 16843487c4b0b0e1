"""Experiment: forward / dgrad L-op launches with the 3xFP16 split (option mma_f16) against fp64 conv on the GPU,
with tiny-magnitude and wide-range inputs (the cases fp16's exponent range makes interesting)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import face_generator_b200 as fg
from face_generator_b200.lib import _ptr

F = torch.nn.functional
dev = lambda a: torch.tensor(a, dtype=torch.float64, device="cuda")
def rel(a, b):
    b = b.cpu().numpy()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))

f = lambda a: np.ascontiguousarray(a, np.float32)
for (N, Cin, H, Cout, k) in [(256, 64, 16, 128, 3), (256, 128, 8, 256, 3), (256, 256, 4, 512, 3), (130, 64, 16, 128, 3), (64, 256, 32, 128, 5)]:
    rng = np.random.default_rng(N + Cin)
    w, b = f(rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)), f(rng.standard_normal(Cout))
    for case, xs, dys in (("unit", 1.0, 1.0), ("tiny-grad", 1.0, 1e-9), ("wide", 3.0, 1e-4)):
        x = f(rng.standard_normal((N, Cin, H, H)) * xs * np.exp(rng.standard_normal((N, Cin, H, H)) * (2.0 if case == "wide" else 0.0)))
        dy = f(rng.standard_normal((N, Cout, H, H)) * dys)
        xt, wt, dyt = dev(x), dev(w), dev(dy)
        ry = F.conv2d(xt, wt, dev(b), padding=k // 2)
        rdx = torch.nn.grad.conv2d_input(xt.shape, wt, dyt, padding=k // 2)
        rdw = torch.nn.grad.conv2d_weight(xt, wt.shape, dyt, padding=k // 2)
        for mode in (0, 1):
            ctx = fg.Context(0, max_batch=8, channels=3)
            ctx.set_option("mma_f16", mode)
            lib, h = ctx.lib, ctx.h
            y, dx = np.empty((N, Cout, H, H), np.float32), np.empty_like(x)
            for rep in range(2):
                ctx.timing_enable(True) if hasattr(ctx, "timing_enable") else None
                assert lib.fg_conv2d_forward(h, _ptr(x), _ptr(w), _ptr(b), _ptr(y), N, Cin, H, H, Cout, k) == 0, lib.fg_last_error()
                assert lib.fg_conv2d_backward_data(h, _ptr(dy), _ptr(w), _ptr(dx), N, Cin, H, H, Cout, k) == 0, lib.fg_last_error()
                dw = np.zeros_like(w)
                assert lib.fg_conv2d_backward_filter(h, _ptr(x), _ptr(dy), _ptr(dw), None, N, Cin, H, H, Cout, k) == 0, lib.fg_last_error()
            print("N%d Cin%d H%d Cout%d k%d %-9s f16=%d  fwd %.2e  dgrad %.2e  wgrad %.2e" % (N, Cin, H, Cout, k, case, mode, rel(y, ry), rel(dx, rdx), rel(dw, rdw)), flush=True)
            ctx.close()
