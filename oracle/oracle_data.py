"""TEST INFRASTRUCTURE ONLY.  numpy restatement of the input side the device dataset replaces:
image.load(path, nbChannels, "float") (= byte/255, rgb2y for nbChannels = 1) followed by image.scale(img, w, h)
(dataset.lua:80-117).  `image` is an un-pinned third-party rock that is absent here; the scaling rule below follows
its generic/image.c (Main_scaleLinear_rowcol, two separable passes: width first, then height).  PARITY UNPINNED."""
import numpy as np


def _axis_matrix(src_len, dst_len):
    """W [dst_len][src_len] in float32 arithmetic exactly as the C loop computes its weights (incl. the division)."""
    W = np.zeros((dst_len, src_len), np.float64)
    f32 = np.float32
    if dst_len > src_len:
        if src_len == 1:
            W[:, 0] = 1.0
            return W
        scale = f32(src_len - 1) / f32(dst_len - 1)
        for di in range(dst_len - 1):
            sf = f32(di) * scale
            si = int(sf)
            fr = f32(sf - f32(si))
            W[di, si] += float(f32(1) - fr)
            W[di, si + 1] += float(fr)
        W[dst_len - 1, src_len - 1] = 1.0
    elif dst_len < src_len:
        scale = f32(src_len) / f32(dst_len)
        si0_i, si0_f = 0, f32(0)
        for di in range(dst_len):
            s1 = f32(di + 1) * scale
            si1_i = int(s1)
            si1_f = f32(s1 - f32(si1_i))
            w = {si0_i: float(f32(1) - si0_f)}
            n = f32(1) - si0_f
            for si in range(si0_i + 1, si1_i):
                w[si] = w.get(si, 0.0) + 1.0
                n = f32(n + f32(1))
            if si1_i < src_len:
                w[si1_i] = w.get(si1_i, 0.0) + float(si1_f)
                n = f32(n + si1_f)
            for k, v in w.items():
                W[di, k] = v / float(n)
            si0_i, si0_f = si1_i, si1_f
    else:
        W[np.arange(dst_len), np.arange(dst_len)] = 1.0
    return W


def load_float(images_u8, nb_channels):
    """image.load(..., nbChannels, 'float') on decoded bytes [N][Cs][H][W]."""
    x = images_u8.astype(np.float64) / 255.0
    if nb_channels == 1 and x.shape[1] == 3:  # image.rgb2y
        x = (0.299 * x[:, 0] + 0.587 * x[:, 1] + 0.114 * x[:, 2])[:, None]
    assert x.shape[1] == nb_channels
    return x


def scale(x, width, height):
    """image.scale(x, width, height), default 'bilinear' mode, on [N][C][H][W] float64."""
    Wx = _axis_matrix(x.shape[3], width)
    Wy = _axis_matrix(x.shape[2], height)
    t = np.einsum("nchw,xw->nchx", x, Wx)
    return np.einsum("nchx,yh->ncyx", t, Wy)


def gather(images_u8, indices, nb_channels, size=32):
    return scale(load_float(images_u8[np.asarray(indices)], nb_channels), size, size)
