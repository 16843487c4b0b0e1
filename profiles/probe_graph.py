"""Experiment: how much of a step is launch overhead?  Captures fg_train_step (device-resident inputs, no stats) into a
CUDA graph through the runtime API and compares graph replay with eager launches."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import face_generator_b200 as fg
from face_generator_b200 import layouts as LY
from face_generator_b200.lib import NET_D, NET_G, _check

rt = C.CDLL("libcudart.so.12")
B, Cc = 256, 3
ctx = fg.Context(0, max_batch=B, channels=Cc)
rng = np.random.default_rng(1)
ctx.set_params(NET_G, LY.trained_like_init(LY.G_layout(Cc), rng))
ctx.set_params(NET_D, LY.trained_like_init(LY.D_layout(Cc), rng, 1.4))
real = rng.random((B // 2, Cc, 32, 32)).astype(np.float32)
nD = rng.uniform(-1, 1, (B // 2, 100)).astype(np.float32)
nG = rng.uniform(-1, 1, (B, 100)).astype(np.float32)
dev = [ctx.dev_array(a) for a in (real, nD, nG)]
hyper = fg.hyper_default()
stream = C.c_void_p()
assert rt.cudaStreamCreateWithFlags(C.byref(stream), 1) == 0
_check(ctx.lib.fg_set_stream(ctx.h, stream), "fg_set_stream")
def step(seed):
    ctx.train_step(hyper, B, dev[0], dev[1], dev[2], None, None, seed, want_stats=False)
for i in range(5):
    step(i)
ctx.sync()
def timed(fn, k):
    ctx.sync(); ctx.event_record(0)
    for i in range(k): fn(i)
    ctx.event_record(1); ctx.sync()
    return ctx.event_elapsed_ms(0, 1) / k
print("eager ms/step", timed(step, 30))
graph, gexec = C.c_void_p(), C.c_void_p()
r = rt.cudaStreamBeginCapture(stream, 0)
assert r == 0, r
step(7)
r = rt.cudaStreamEndCapture(stream, C.byref(graph))
assert r == 0, ("end capture", r)
rt.cudaGraphInstantiate.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_ulonglong]
r = rt.cudaGraphInstantiate(C.byref(gexec), graph, 0)
assert r == 0, ("instantiate", r)
def replay(i):
    assert rt.cudaGraphLaunch(gexec, stream) == 0
for i in range(3): replay(i)
print("graph ms/step", timed(replay, 30))
