"""Host-side mirror of dataset.lua for the device-resident path: decoded uint8 images live on the GPU, the batch
(`inputs[i] = dataset[math.random(dataset:size())]`, adversarial.lua:244-249) is assembled by one kernel.

    ds = DeviceDataset(ctx, images_u8)            # [N][Cs][Hs][Ws] uint8, e.g. the 64x64 faces of dataset.lua:10
    real = ds.gather(indices)                     # == image.scale(image.load(...), 32, 32) for those images
    stats = ds.train_step(hyper, B, seed)         # adversarial.lua loop body with no host->device traffic
"""
import ctypes as C

import numpy as np

from .lib import Context, StepStats, _check


class DeviceDataset:
    def __init__(self, ctx: Context, images_u8, chunk=8192):
        images_u8 = np.ascontiguousarray(images_u8, np.uint8)
        assert images_u8.ndim == 4, "[N][Cs][Hs][Ws] uint8"
        N, Cs, Hs, Ws = images_u8.shape
        self.ctx, self.lib, self.N = ctx, ctx.lib, N
        h = C.c_void_p()
        _check(self.lib.fg_dataset_create(ctx.h, N, Cs, Hs, Ws, C.byref(h)), "fg_dataset_create")
        self.h = h
        for s in range(0, N, chunk):  # dataset.loadImages(startAt, count) granularity
            part = images_u8[s:s + chunk]
            _check(self.lib.fg_dataset_upload(self.h, s, part.shape[0], part.ctypes.data_as(C.c_void_p)), "fg_dataset_upload")

    def close(self):
        if self.h:
            self.lib.fg_dataset_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.close()
        except Exception:
            pass

    def size(self):
        return int(self.lib.fg_dataset_size(self.h))

    def gather(self, indices):
        idx = np.ascontiguousarray(indices, np.int32)
        out = np.empty((idx.size, self.ctx.C, 32, 32), np.float32)
        _check(self.lib.fg_dataset_gather(self.h, idx.ctypes.data_as(C.c_void_p), idx.size, out.ctypes.data_as(C.c_void_p)),
               "fg_dataset_gather")
        return out

    def draw(self, seed, B):
        idx = np.empty(B, np.int32)
        _check(self.lib.fg_dataset_draw(self.h, seed, B, idx.ctypes.data_as(C.c_void_p)), "fg_dataset_draw")
        return idx

    def train_step(self, hyper, B, seed, want_stats=True):
        st = StepStats() if want_stats else None
        _check(self.lib.fg_train_step_dataset(self.ctx.h, self.h, C.byref(hyper), B, seed,
                                              C.byref(st) if st is not None else None), "fg_train_step_dataset")
        if st is None:
            return None
        return dict(loss_D=st.loss_D, loss_G=st.loss_G, conf=list(st.conf), trained_D=st.trained_D, t_D=st.t_D,
                    t_G=st.t_G, acc_D=st.acc_D)


def noise_uniform(ctx: Context, seed, shape):
    """NN_UTILS.createNoiseInputs drawn on the device (the stream fg_train_step_dataset uses)."""
    out = np.empty(shape, np.float32)
    _check(ctx.lib.fg_noise_uniform(ctx.h, seed, out.size, out.ctypes.data_as(C.c_void_p)), "fg_noise_uniform")
    return out
