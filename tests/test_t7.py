"""CPU: Torch7 checkpoint reader / writer behind the C ABI (fg_t7_*), SURVEY.md section 8(f).1.

No Torch7 exists in this image, so the files are produced / parsed here by a second, independent implementation of
the torch7 File.lua object format (struct-packing below) -- PARITY UNPINNED for this format as well.  The module
trees mimic what the reference saves: torch.save(filename, {D=MODEL_D, G=MODEL_G, opt=OPT, epoch=EPOCH})
(adversarial.lua:328) with the nets in CUDA mode, i.e. nn.Sequential{nn.Copy, nn.Sequential{...}, nn.Copy}
(utils/nn_utils.lua:328-363) whose parameters are views into ONE flat storage after getParameters() (train.lua:151)."""
import struct

import numpy as np
import pytest

from face_generator_b200 import layouts as LY
from face_generator_b200.checkpoint import T7File, T7Writer
from face_generator_b200.lib import FGError


# ------------------------------------------------------------------ independent writer (test side)
class Obj:
    def __init__(self, cls, fields):
        self.cls, self.fields = cls, fields


class Tensor:
    def __init__(self, storage, size, stride=None, offset=0, cls="torch.FloatTensor"):
        self.storage, self.size, self.offset, self.cls = storage, list(size), offset, cls
        if stride is None:
            stride, s = [], 1
            for d in reversed(self.size):
                stride.insert(0, s)
                s *= d
        self.stride = list(stride)


class Storage:
    def __init__(self, data, cls="torch.FloatStorage"):
        self.data, self.cls = data, cls


class W:
    def __init__(self):
        self.buf, self.ids = bytearray(), {}

    def i32(self, v):
        self.buf += struct.pack("<i", v)

    def i64(self, v):
        self.buf += struct.pack("<q", v)

    def s(self, v):
        b = v.encode()
        self.i32(len(b))
        self.buf += b

    def ref(self, o, typ):
        self.i32(typ)
        if id(o) in self.ids:
            self.i32(self.ids[id(o)])
            return True
        self.ids[id(o)] = len(self.ids) + 1
        self.i32(self.ids[id(o)])
        return False

    def obj(self, o):
        if o is None:
            self.i32(0)
        elif isinstance(o, bool):
            self.i32(5)
            self.i32(1 if o else 0)
        elif isinstance(o, (int, float)):
            self.i32(1)
            self.buf += struct.pack("<d", float(o))
        elif isinstance(o, str):
            self.i32(2)
            self.s(o)
        elif isinstance(o, dict):
            if self.ref(o, 3):
                return
            self.i32(len(o))
            for k, v in o.items():
                self.obj(k)
                self.obj(v)
        elif isinstance(o, Tensor):
            if self.ref(o, 4):
                return
            self.s("V 1")
            self.s(o.cls)
            self.i32(len(o.size))
            for d in o.size:
                self.i64(d)
            for d in o.stride:
                self.i64(d)
            self.i64(o.offset + 1)
            self.obj(o.storage)
        elif isinstance(o, Storage):
            if self.ref(o, 4):
                return
            self.s("V 1")
            self.s(o.cls)
            self.i64(o.data.size)
            self.buf += o.data.tobytes()
        elif isinstance(o, Obj):
            if self.ref(o, 4):
                return
            self.s("V 1")
            self.s(o.cls)
            self.obj(o.fields)
        else:
            raise TypeError(type(o))


def seq(*mods):
    return Obj("nn.Sequential", {"modules": {i + 1: m for i, m in enumerate(mods)}, "train": True,
                                 "output": Tensor(None, []), "gradInput": Tensor(None, [])})


def cuda_net(flat, layout, classes, tensor_cls="torch.CudaTensor", storage_cls="torch.CudaStorage", bn=None):
    """nn.Sequential{Copy, Sequential{layers...}, Copy}; every weight/bias is a view into the one flat storage."""
    st = Storage(flat, storage_cls)
    gst = Storage(np.zeros_like(flat), storage_cls)
    mods, items = [], list(layout.items())
    i = 0
    for cls, nparam in classes:
        fields = {"train": True}
        for j in range(nparam):
            name, (off, shape) = items[i]
            key = "weight" if j == 0 else "bias"
            fields[key] = Tensor(st, shape, offset=off, cls=tensor_cls)
            fields["grad" + key.capitalize()] = Tensor(gst, shape, offset=off, cls=tensor_cls)
            i += 1
        if "BatchNormalization" in cls and bn is not None:
            fields.update(bn.pop(0))
        mods.append(Obj(cls, fields))
    assert i == len(items)
    copy = lambda a, b: Obj("nn.Copy", {"intype": a, "outtype": b, "train": True})
    return seq(copy("torch.FloatTensor", "torch.CudaTensor"), seq(*mods), copy("torch.CudaTensor", "torch.FloatTensor"))


G_CLASSES = [("nn.Linear", 2), ("nn.View", 0), ("nn.PReLU", 1), ("nn.SpatialUpSamplingNearest", 0),
             ("cudnn.SpatialConvolution", 2), ("nn.SpatialBatchNormalization", 2), ("nn.PReLU", 1),
             ("nn.SpatialUpSamplingNearest", 0), ("cudnn.SpatialConvolution", 2), ("nn.SpatialBatchNormalization", 2),
             ("nn.PReLU", 1), ("cudnn.SpatialConvolution", 2), ("nn.Sigmoid", 0)]
D_CLASSES = ([("nn.SpatialConvolution", 2), ("nn.PReLU", 1), ("nn.SpatialDropout", 0), ("nn.SpatialAveragePooling", 0)] * 4 +
             [("nn.View", 0), ("nn.Linear", 2), ("nn.PReLU", 1), ("nn.Dropout", 0), ("nn.Linear", 2), ("nn.PReLU", 1),
              ("nn.Dropout", 0), ("nn.Linear", 2), ("nn.Sigmoid", 0)])


def fstore(a):
    return Storage(np.ascontiguousarray(a, np.float32))


def write_reference_like(path, C=3, seed=1, bn_style="var"):
    rng = np.random.default_rng(seed)
    (gl, ng), (dl, nd) = LY.G_layout(C), LY.D_layout(C)
    PG, PD = rng.standard_normal(ng).astype(np.float32), rng.standard_normal(nd).astype(np.float32)
    rm1, rv1 = rng.standard_normal(256).astype(np.float32), rng.uniform(0.5, 2, 256).astype(np.float32)
    rm2, rv2 = rng.standard_normal(128).astype(np.float32), rng.uniform(0.5, 2, 128).astype(np.float32)
    t1 = lambda a: Tensor(fstore(a), [a.size], cls="torch.CudaTensor")
    if bn_style == "var":
        bn = [{"running_mean": t1(rm1), "running_var": t1(rv1), "eps": 1e-5, "momentum": 0.1},
              {"running_mean": t1(rm2), "running_var": t1(rv2), "eps": 1e-5, "momentum": 0.1}]
    else:  # 2015 nn: running_std = 1/sqrt(var + eps)
        bn = [{"running_mean": t1(rm1), "running_std": t1(1 / np.sqrt(rv1 + 1e-5)), "eps": 1e-5},
              {"running_mean": t1(rm2), "running_std": t1(1 / np.sqrt(rv2 + 1e-5)), "eps": 1e-5}]
    root = {"G": cuda_net(PG, gl, G_CLASSES, bn=bn), "D": cuda_net(PD, dl, D_CLASSES),
            "opt": {"batchSize": 32, "save": "logs", "grayscale": C == 1, "D_L2": 1e-4}, "epoch": 7}
    w = W()
    w.obj(root)
    open(path, "wb").write(bytes(w.buf))
    return PG, PD, np.concatenate([rm1, rv1, rm2, rv2])


# ------------------------------------------------------------------ tests
@pytest.mark.parametrize("C,bn_style", [(3, "var"), (1, "std")])
def test_reads_reference_style_checkpoint(tmp_path, C, bn_style):
    p = tmp_path / "adversarial.net"
    PG, PD, bn = write_reference_like(p, C=C, bn_style=bn_style)
    with T7File(p) as f:
        assert f.kind("G") == "object" and f.string("G") == "nn.Sequential" and f.kind("nope") is None
        assert f.number("epoch") == 7 and f.number("opt.batchSize") == 32 and f.string("opt.save") == "logs"
        assert f.number("opt.grayscale") == (1.0 if C == 1 else 0.0)
        np.testing.assert_array_equal(f.net_params("G"), PG)  # getParameters() order, views into one storage
        np.testing.assert_array_equal(f.net_params("D"), PD)
        got_bn = f.net_bn_state("G")
        assert got_bn.size == 768
        np.testing.assert_allclose(got_bn, bn, rtol=2e-6 if bn_style == "std" else 0)
        assert f.net_bn_state("D").size == 0
        d = f.net_describe("G")
        assert d.startswith("nn.Sequential{nn.Copy,nn.Sequential{nn.Linear,nn.View,nn.PReLU,nn.SpatialUpSamplingNearest,cudnn.Spa")
        assert d.endswith("nn.Sigmoid},nn.Copy}")
        # a single module / a single tensor by dotted path; numeric segments index the array part
        assert f.string("G.modules.2.modules.1") == "nn.Linear"
        w = f.tensor("G.modules.2.modules.1.weight")
        assert w.shape == (8192, 100)
        np.testing.assert_array_equal(w.ravel(), PG[:819200])
        np.testing.assert_array_equal(f.net_params("G.modules.2.modules.5"), PG[LY.G_layout(C)[0]["C1W"][0]:][:819456])


def test_strides_offsets_dtypes_and_shared_storage(tmp_path):
    base = np.arange(24, dtype=np.float64)
    st = Storage(base, "torch.DoubleStorage")
    a = Tensor(st, [2, 3], stride=[1, 2], offset=4, cls="torch.DoubleTensor")  # transposed view at offset 4
    lng = Tensor(Storage(np.array([5, -7, 9], np.int64), "torch.LongStorage"), [3], cls="torch.LongTensor")
    byt = Tensor(Storage(np.array([1, 255], np.uint8), "torch.ByteStorage"), [2], cls="torch.ByteTensor")
    shared = {"x": 1.5}
    root = {"a": a, "again": a, "lng": lng, "byt": byt, "t1": shared, "t2": shared, "flag": True, "none": None,
            "empty": Tensor(None, []), 1: "first", 2: "second"}
    w = W()
    w.obj(root)
    p = tmp_path / "misc.t7"
    p.write_bytes(bytes(w.buf))
    with T7File(p) as f:
        exp = np.array([[base[4 + i * 1 + j * 2] for j in range(3)] for i in range(2)], np.float32)
        np.testing.assert_array_equal(f.tensor("a"), exp)
        np.testing.assert_array_equal(f.tensor("again"), exp)  # second occurrence is a reference to the first
        np.testing.assert_array_equal(f.tensor("lng"), [5, -7, 9])
        np.testing.assert_array_equal(f.tensor("byt"), [1, 255])
        assert f.number("t1.x") == 1.5 and f.number("t2.x") == 1.5 and f.number("flag") == 1.0
        assert f.kind("none") == "nil" and f.tensor("empty").size == 0
        assert f.string("1") == "first" and f.string("2") == "second"
        with pytest.raises(FGError):
            f.number("a")
        with pytest.raises(FGError):
            f.tensor("t1")


def test_bad_files_fail_loudly(tmp_path):
    p = tmp_path / "adversarial.net"
    write_reference_like(p, C=1)
    data = p.read_bytes()
    (tmp_path / "trunc.net").write_bytes(data[:len(data) // 2])
    with pytest.raises(FGError):
        T7File(tmp_path / "trunc.net")
    (tmp_path / "ascii.net").write_text("3\n1\n2\n")  # torch.save(..., 'ascii') is not supported
    with pytest.raises(FGError):
        T7File(tmp_path / "ascii.net")
    with pytest.raises(FGError):
        T7File(tmp_path / "missing.net")
    # length fields larger than the file are rejected before anything is allocated (nothing may throw or exhaust
    # memory across the C ABI)
    ps = lambda x: struct.pack("<i", len(x)) + x
    hostile = {
        "bigstorage": struct.pack("<ii", 4, 1) + ps(b"V 1") + ps(b"torch.FloatStorage") + struct.pack("<q", 1 << 35),
        "bigstring": struct.pack("<ii", 2, 1 << 27),
        "bigtable": struct.pack("<iii", 3, 1, 1 << 25),
    }
    for name, blob in hostile.items():
        (tmp_path / name).write_bytes(blob)
        with pytest.raises(FGError):
            T7File(tmp_path / name)
    huge = (struct.pack("<iii", 3, 1, 1) + struct.pack("<i", 2) + ps(b"a") + struct.pack("<ii", 4, 2) + ps(b"V 1") +
            ps(b"torch.FloatTensor") + struct.pack("<i", 2) + struct.pack("<qq", 1 << 40, 1 << 40) + struct.pack("<qq", 1, 1) +
            struct.pack("<q", 1) + struct.pack("<ii", 4, 3) + ps(b"V 1") + ps(b"torch.FloatStorage") + struct.pack("<q", 2) +
            struct.pack("<ff", 1, 2))
    (tmp_path / "hugetensor").write_bytes(huge)
    with T7File(tmp_path / "hugetensor") as f:
        with pytest.raises(FGError):
            f.tensor("a")
    # a tensor that points outside its storage must not be read
    w = W()
    w.obj({"bad": Tensor(fstore(np.zeros(4)), [8])})
    (tmp_path / "oob.t7").write_bytes(bytes(w.buf))
    with T7File(tmp_path / "oob.t7") as f:
        with pytest.raises(FGError):
            f.tensor("bad")
    # an "expanded" (stride-0) weight of 2^31 logical elements over a 1-element storage: a ~300-byte file must not
    # turn into a multi-GiB flatten, neither through the tensor nor through the net accessors (count-only calls too)
    w = W()
    big = Tensor(fstore(np.zeros(1)), [1 << 16, 1 << 15], stride=[0, 0])
    w.obj({"t": big, "G": seq(Obj("nn.Linear", {"weight": big, "bias": big, "running_mean": big, "running_var": big}))})
    (tmp_path / "expanded.t7").write_bytes(bytes(w.buf))
    assert len(w.buf) < 1000
    with T7File(tmp_path / "expanded.t7") as f:
        for call in (lambda: f.tensor("t"), lambda: f.net_params("G"), lambda: f.net_bn_state("G")):
            with pytest.raises(FGError):
                call()
    # the object memo allows a `modules` table to contain its own parent: the walkers must refuse the cycle ...
    w = W()
    inner = seq(Obj("nn.Linear", {"weight": Tensor(fstore(np.ones(2)), [2])}))
    inner.fields["modules"][2] = inner
    w.obj({"G": inner})
    (tmp_path / "cycle.t7").write_bytes(bytes(w.buf))
    with T7File(tmp_path / "cycle.t7") as f:
        for call in (lambda: f.net_params("G"), lambda: f.net_bn_state("G"), lambda: f.net_describe("G")):
            with pytest.raises(FGError):
                call()
    # ... and a DAG that repeats one subtree 4^20 times (tiny file, exponential walk) must stop at the visit budget
    w = W()
    node = seq(Obj("nn.Linear", {"weight": Tensor(fstore(np.ones(2)), [2])}))
    for _ in range(20):
        node = seq(node, node, node, node)
    w.obj({"G": node})
    (tmp_path / "fanout.t7").write_bytes(bytes(w.buf))
    assert len(w.buf) < 20000
    import time
    with T7File(tmp_path / "fanout.t7") as f:
        t0 = time.time()
        for call in (lambda: f.net_params("G"), lambda: f.net_describe("G")):
            with pytest.raises(FGError):
                call()
        assert time.time() - t0 < 20


# ------------------------------------------------------------------ independent reader for the writer test
def parse(buf):
    pos, memo = [0], {}

    def rd(fmt):
        v = struct.unpack_from("<" + fmt, buf, pos[0])
        pos[0] += struct.calcsize("<" + fmt)
        return v[0]

    def rs():
        n = rd("i")
        s = buf[pos[0]:pos[0] + n].decode()
        pos[0] += n
        return s

    def obj():
        t = rd("i")
        if t == 0:
            return None
        if t == 1:
            return rd("d")
        if t == 2:
            return rs()
        if t == 5:
            return rd("i") == 1
        idx = rd("i")
        if idx in memo:
            return memo[idx]
        if t == 3:
            out = memo[idx] = {}
            for _ in range(rd("i")):
                k = obj()
                out[k] = obj()
            return out
        assert t == 4
        assert rs() == "V 1"
        cls = rs()
        if cls.endswith("Tensor"):
            nd = rd("i")
            size = [rd("q") for _ in range(nd)]
            stride = [rd("q") for _ in range(nd)]
            off = rd("q")
            out = memo[idx] = dict(cls=cls, size=size, stride=stride, offset=off, storage=obj())
            return out
        n = rd("q")
        data = np.frombuffer(buf, np.float32, n, pos[0]).copy()
        pos[0] += 4 * n
        out = memo[idx] = dict(cls=cls, data=data)
        return out

    root = obj()
    assert pos[0] == len(buf), "trailing bytes"
    return root


def test_writer_produces_stock_torch_layout_and_round_trips(tmp_path):
    rng = np.random.default_rng(3)
    PG = rng.standard_normal(1000).astype(np.float32)
    m = rng.standard_normal((4, 5, 6)).astype(np.float32)
    p = tmp_path / "flat.t7"
    w = T7Writer(p)
    w.add("G", PG)
    w.add("m", m)
    w.add("epoch", 12)
    w.add("note", "hello")
    w.close()
    root = parse(p.read_bytes())
    assert set(root) == {"G", "m", "epoch", "note"} and root["epoch"] == 12.0 and root["note"] == "hello"
    g = root["G"]
    assert g["cls"] == "torch.FloatTensor" and g["size"] == [1000] and g["stride"] == [1] and g["offset"] == 1
    assert g["storage"]["cls"] == "torch.FloatStorage"
    np.testing.assert_array_equal(g["storage"]["data"], PG)
    assert root["m"]["size"] == [4, 5, 6] and root["m"]["stride"] == [30, 6, 1]
    np.testing.assert_array_equal(root["m"]["storage"]["data"], m.ravel())
    with T7File(p) as f:
        np.testing.assert_array_equal(f.tensor("G"), PG)
        np.testing.assert_array_equal(f.tensor("m"), m)
        assert f.number("epoch") == 12 and f.string("note") == "hello"


def test_c2f_checkpoint_params(tmp_path):
    """adversarial_c2f.lua:216 saves the same {D, G, opt, epoch} table for the coarse-to-fine nets."""
    C = 3
    (gl, ng), (dl, nd) = LY.c2f_G_layout(C), LY.c2f_D_layout(C)
    rng = np.random.default_rng(5)
    PG = rng.standard_normal(ng).astype(np.float32)
    g_classes = [("cudnn.SpatialConvolutionUpsample", 2), ("nn.PReLU", 1)] * 4 + [("cudnn.SpatialConvolutionUpsample", 2),
                                                                                   ("nn.View", 0)]
    inner = cuda_net(PG, gl, g_classes)
    root = {"G": seq(Obj("nn.JoinTable", {"dimension": 2, "nInputDims": 2}), *inner.fields["modules"].values()), "epoch": 3}
    w = W()
    w.obj(root)
    p = tmp_path / "adversarial_c2f.net"
    p.write_bytes(bytes(w.buf))
    with T7File(p) as f:
        np.testing.assert_array_equal(f.net_params("G"), PG)
        assert f.net_describe("G").startswith("nn.Sequential{nn.JoinTable,nn.Copy,nn.Sequential{cudnn.SpatialConvolutionUpsample,")


@pytest.mark.gpu
def test_gpu_checkpoint_load_and_resume(tmp_path):
    """sample.lua:247-258 on a reference-style checkpoint, then save / resume with the optimizer state kept."""
    import face_generator_b200 as fg
    from face_generator_b200 import checkpoint as CK
    from face_generator_b200.lib import NET_D, NET_G
    import parity_utils as PU
    B, C = 8, 3
    p = tmp_path / "adversarial.net"
    PG, PD, bn = write_reference_like(p, C=C)
    ctx = fg.Context(0, max_batch=B, channels=C)
    assert CK.load_reference_checkpoint(ctx, p) == 7
    np.testing.assert_array_equal(ctx.get_params(NET_G), PG)
    np.testing.assert_array_equal(ctx.get_params(NET_D), PD)
    np.testing.assert_array_equal(ctx.get_bn_state(), bn)
    # a gray checkpoint does not fit a colour context: refuse instead of loading garbage
    pg = tmp_path / "gray.net"
    write_reference_like(pg, C=1)
    with pytest.raises(fg.FGError):
        CK.load_reference_checkpoint(ctx, pg)
    # train two steps from a sane init, save, resume in a fresh context, and take the same third step in both
    case = PU.make_case(B, C, seed=77)
    ctx.set_params(NET_G, case["PG"])
    ctx.set_params(NET_D, case["PD"])
    hyper = fg.hyper_default()
    args = (hyper, B, case["real"], case["noise_D"], case["noise_G"], case["masks_D"], case["masks_G"])
    for _ in range(2):
        ctx.train_step(*args)
    f = tmp_path / "resume.t7"
    CK.save_flat_checkpoint(ctx, f, epoch=3)
    ctx2 = fg.Context(0, max_batch=B, channels=C)
    assert CK.load_flat_checkpoint(ctx2, f) == 3
    for net in (NET_G, NET_D):
        np.testing.assert_array_equal(ctx2.get_params(net), ctx.get_params(net))
        m1, v1, t1 = ctx.get_adam_state(net)
        m2, v2, t2 = ctx2.get_adam_state(net)
        assert t1 == t2 == 2
        np.testing.assert_array_equal(m1, m2)
        np.testing.assert_array_equal(v1, v2)
    s1, s2 = ctx.train_step(*args), ctx2.train_step(*args)
    assert s1["t_D"] == s2["t_D"] == 3
    # loss_D only depends on the (identical) restored state; loss_G and the new parameters come after D's update,
    # where split-K atomics can flip a PReLU branch between the two runs
    assert abs(s1["loss_D"] - s2["loss_D"]) < 1e-5 and abs(s1["loss_G"] - s2["loss_G"]) < 2e-3
    assert PU.relerr(ctx2.get_params(NET_D), ctx.get_params(NET_D)) < 5e-3
    ctx.close()
    ctx2.close()
