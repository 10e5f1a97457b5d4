"""Torch7 binary serialisation (torch.save / torch.load) -- reader, writer and the nn-module-tree walk that getParameters() does.

SURVEY.md section 8(f) row F2.  The reference writes `torch.save(filename, {D = MODEL_D, G = MODEL_G, opt = OPT, plot_data = ...,
epoch = EPOCH, normalize_mean = ..., normalize_std = ...})` (/root/reference/train.lua:252-261), reads it back with torch.load
(:127-137) and loads V the same way (:119-123).  The format belongs to torch7's File.lua / Tensor.c / Storage.c, an un-vendored
dependency of the reference (no copy under /root/reference, no .net fixture either), so it is restated here from its published
definition and is PARITY-UNPINNED: the tests round-trip this writer through this reader and check hand-assembled byte strings, they
cannot check a file Torch7 itself produced.

Format (binary mode, native little-endian; int = 4 bytes, long = 8 bytes, double = 8 bytes):
  object   := int type, then
     0 nil
     1 number   : double
     2 string   : int length, bytes
     5 boolean  : int 0/1
     3 table    : int index; if the index was seen before, nothing more (a back-reference); else int npairs, npairs x (object key, object value)
     4 torch    : int index; back-reference as above; else string "V <n>" (class version), string class name, then the class's
                  own `write`: tensors and storages below, every other class (all nn modules) writes ONE object: the table of its fields
     6,7,8 function : int index; back-reference as above; else int length, bytes (a string.dump), object upvalues
  tensor   := int ndim, long size[ndim], long stride[ndim], long storageOffset (1-based), object storage (or nil)
  storage  := long n, n raw elements
getParameters() (train.lua:184-185) leaves every weight / bias a VIEW into one flat storage, so a saved network holds that storage
once and the other tensors as back-references with different offsets; the reader honours offsets and strides.
"""
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN, TYPE_FUNCTION, LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = range(9)

_ELEM = {"Float": np.float32, "Double": np.float64, "Long": np.int64, "Int": np.int32, "Short": np.int16, "Char": np.int8,
         "Byte": np.uint8, "Half": np.float16, "Cuda": np.float32, "CudaDouble": np.float64, "CudaLong": np.int64,
         "CudaInt": np.int32, "CudaByte": np.uint8, "CudaHalf": np.float16}
_NAME = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.int64): "Long", np.dtype(np.int32): "Int",
         np.dtype(np.uint8): "Byte", np.dtype(np.int16): "Short", np.dtype(np.int8): "Char"}


class T7Error(ValueError):
    pass


class T7Object:
    """A torch class instance that is not a tensor or storage (every nn module): class name + field table."""

    def __init__(self, typename, fields=None, version=1):
        self.typename, self.fields, self.version = typename, ({} if fields is None else fields), version

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def __repr__(self):
        return "T7Object(%s)" % self.typename


class T7Function:
    def __init__(self, dumped, upvalues):
        self.dumped, self.upvalues = dumped, upvalues


class T7Storage:
    """A torch.*Storage; `data` is a 1-D numpy array.  Identity matters: tensors sharing it are written as back-references."""

    def __init__(self, data, typename=None):
        self.data = np.ascontiguousarray(data).reshape(-1)
        self.typename = typename or "torch.%sStorage" % _NAME[self.data.dtype]


class T7Tensor:
    """A strided view (size, stride, offset in elements) into a T7Storage; array() materialises it."""

    def __init__(self, storage, size, stride=None, offset=0, typename=None):
        self.storage, self.size, self.offset = storage, tuple(int(s) for s in size), int(offset)
        if stride is None:
            stride, acc = [], 1
            for s in reversed(self.size):
                stride.append(acc); acc *= s
            stride = tuple(reversed(stride))
        self.stride = tuple(int(s) for s in stride)
        self.typename = typename or (storage.typename.replace("Storage", "Tensor") if storage is not None else "torch.FloatTensor")

    @staticmethod
    def of(a, typename=None):
        a = np.ascontiguousarray(a)
        return T7Tensor(T7Storage(a), a.shape, typename=typename)

    def array(self):
        if self.storage is None or not self.size:
            return np.zeros(self.size if self.size else (0,), np.float32)
        d = self.storage.data
        need = self.offset + sum((s - 1) * st for s, st in zip(self.size, self.stride)) + 1 if all(self.size) else 0
        if need > d.size or self.offset < 0:
            raise T7Error("tensor view [%s] x strides [%s] + %d exceeds its storage of %d elements" % (self.size, self.stride, self.offset, d.size))
        v = np.lib.stride_tricks.as_strided(d[self.offset:], shape=self.size, strides=tuple(s * d.itemsize for s in self.stride), writeable=False)
        return np.array(v)


# ---------------------------------------------------------------------------------------------------------------- reader
class Reader:
    def __init__(self, buf):
        self.b, self.o, self.objects = memoryview(buf), 0, {}

    def _take(self, n):
        if n < 0 or self.o + n > len(self.b):
            raise T7Error("truncated Torch7 file: need %d bytes at offset %d of %d" % (n, self.o, len(self.b)))
        v = self.b[self.o:self.o + n]; self.o += n
        return v

    def int(self):
        return struct.unpack("<i", self._take(4))[0]

    def long(self):
        return struct.unpack("<q", self._take(8))[0]

    def string(self):
        return bytes(self._take(self.int())).decode("latin-1")

    def read(self):
        t = self.int()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = struct.unpack("<d", self._take(8))[0]
            return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_BOOLEAN:
            return self.int() != 0
        if t not in (TYPE_TABLE, TYPE_TORCH, TYPE_FUNCTION, LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION):
            raise T7Error("unknown Torch7 type tag %d at offset %d" % (t, self.o - 4))
        index = self.int()
        if index in self.objects:
            return self.objects[index]
        if t == TYPE_TABLE:
            n = self.int()
            tab = self.objects[index] = {}
            for _ in range(n):
                k = self.read()
                tab[k] = self.read()
            return tab
        if t != TYPE_TORCH:
            f = self.objects[index] = T7Function(bytes(self._take(self.int())), None)
            f.upvalues = self.read()
            return f
        version = self.string()
        if version.startswith("V "):
            vnum, cls = int(version[2:]), self.string()
        else:
            vnum, cls = 0, version                     # files older than class versioning carry the class name first
        if cls.startswith("torch.") and cls.endswith("Storage"):
            dt = self._dtype(cls[6:-7], cls)
            n = self.long()
            s = self.objects[index] = T7Storage(np.frombuffer(bytes(self._take(n * np.dtype(dt).itemsize)), dt), cls)
            return s
        if cls.startswith("torch.") and cls.endswith("Tensor"):
            self._dtype(cls[6:-6], cls)
            nd = self.int()
            size = [self.long() for _ in range(nd)]
            stride = [self.long() for _ in range(nd)]
            off = self.long() - 1
            st = self.read()
            if st is not None and not isinstance(st, T7Storage):
                raise T7Error("%s holds a %r where its storage belongs" % (cls, st))
            t7 = self.objects[index] = T7Tensor(st, size, stride, max(off, 0), cls)
            return t7
        obj = self.objects[index] = T7Object(cls, None, vnum)
        fields = self.read()
        if not isinstance(fields, dict):
            raise T7Error("%s: expected its field table, found %r" % (cls, fields))
        obj.fields = fields
        return obj

    @staticmethod
    def _dtype(name, cls):
        if name not in _ELEM:
            raise T7Error("unsupported Torch7 class %s" % cls)
        return _ELEM[name]


def load(path):
    """torch.load(path) for binary files: tables -> dict, numbers -> int/float, tensors -> T7Tensor, modules -> T7Object."""
    with open(path, "rb") as f:
        buf = f.read()
    r = Reader(buf)
    obj = r.read()
    if r.o != len(buf):
        raise T7Error("%d trailing bytes after the top-level object" % (len(buf) - r.o))
    return obj


# ---------------------------------------------------------------------------------------------------------------- writer
class Writer:
    def __init__(self):
        self.out, self.index, self.keep = [], {}, []

    def int(self, v):
        self.out.append(struct.pack("<i", int(v)))

    def long(self, v):
        self.out.append(struct.pack("<q", int(v)))

    def string(self, s):
        b = s.encode("latin-1"); self.int(len(b)); self.out.append(b)

    def _ref(self, tag, obj):
        """Writes the tag and the object's index; True when the object was written before (back-reference only)."""
        self.int(tag)
        seen = id(obj) in self.index
        if not seen:
            self.index[id(obj)] = len(self.index) + 1
            self.keep.append(obj)                     # ids stay unique while the writer lives
        self.int(self.index[id(obj)])
        return seen

    def write(self, v):
        if v is None:
            self.int(TYPE_NIL)
        elif isinstance(v, (bool, np.bool_)):
            self.int(TYPE_BOOLEAN); self.int(1 if v else 0)
        elif isinstance(v, (int, float, np.integer, np.floating)):
            self.int(TYPE_NUMBER); self.out.append(struct.pack("<d", float(v)))
        elif isinstance(v, str):
            self.int(TYPE_STRING); self.string(v)
        elif isinstance(v, np.ndarray):
            self.write(T7Tensor.of(v))
        elif isinstance(v, (list, tuple)):
            self.write({i + 1: x for i, x in enumerate(v)})       # Lua arrays are 1-based tables
        elif isinstance(v, dict):
            if not self._ref(TYPE_TABLE, v):
                self.int(len(v))
                for k, x in v.items():
                    self.write(k); self.write(x)
        elif isinstance(v, T7Storage):
            if not self._ref(TYPE_TORCH, v):
                self.string("V 1"); self.string(v.typename)
                self.long(v.data.size); self.out.append(v.data.tobytes())
        elif isinstance(v, T7Tensor):
            if not self._ref(TYPE_TORCH, v):
                self.string("V 1"); self.string(v.typename)
                self.int(len(v.size))
                for s in v.size:
                    self.long(s)
                for s in v.stride:
                    self.long(s)
                self.long(v.offset + 1)
                self.write(v.storage)
        elif isinstance(v, T7Object):
            if not self._ref(TYPE_TORCH, v):
                self.string("V %d" % v.version); self.string(v.typename)
                self.write(v.fields)
        elif isinstance(v, T7Function):
            if not self._ref(TYPE_RECUR_FUNCTION, v):
                self.int(len(v.dumped)); self.out.append(v.dumped); self.write(v.upvalues)
        else:
            raise T7Error("cannot serialise %r" % type(v))

    def bytes(self):
        return b"".join(self.out)


def dumps(obj):
    w = Writer(); w.write(obj)
    return w.bytes()


def save(path, obj):
    with open(path, "wb") as f:
        f.write(dumps(obj))


# ------------------------------------------------------------------------------- nn module trees <-> flat parameter vectors
def _children(mod):
    m = mod.get("modules")
    if not isinstance(m, dict):
        return None
    return [m[i] for i in sorted(k for k in m if isinstance(k, int))]


def parameters(mod):
    """Module:parameters() the way nn defines it: containers concatenate their children's in order; a leaf gives {weight, bias}."""
    kids = _children(mod)
    if kids is not None:
        return [p for k in kids for p in parameters(k)]
    out = []
    for name in ("weight", "bias"):
        t = mod.get(name)
        if isinstance(t, T7Tensor) and t.storage is not None and int(np.prod(t.size)) > 0:
            out.append(t)
    return out


def flat_parameters(mod):
    """The vector MODEL:getParameters() returns (train.lua:184-185), as float32."""
    ps = parameters(mod)
    return np.concatenate([p.array().astype(np.float32).reshape(-1) for p in ps]) if ps else np.zeros(0, np.float32)


def bn_running(mod, eps=1e-5):
    """[mean, var] of every (Spatial)BatchNormalization in traversal order.  nn before 2016 stored running_std = 1/sqrt(var + eps)."""
    kids = _children(mod)
    if kids is not None:
        parts = [bn_running(k, eps) for k in kids]
        parts = [p for p in parts if p.size]
        return np.concatenate(parts) if parts else np.zeros(0, np.float32)
    if not mod.typename.endswith("BatchNormalization"):
        return np.zeros(0, np.float32)
    mean = mod["running_mean"].array().astype(np.float32)
    if isinstance(mod.get("running_var"), T7Tensor):
        var = mod["running_var"].array().astype(np.float32)
    else:
        std = mod["running_std"].array().astype(np.float64)
        var = (1.0 / (std * std) - float(mod.get("eps", eps))).astype(np.float32)
    return np.concatenate([mean, var])


# ---- module trees shaped like models.lua's constructors, filled from flat vectors (what torch.save would hold after getParameters())
class _Flat:
    def __init__(self, params):
        self.P = T7Storage(np.ascontiguousarray(params, np.float32))
        self.G = T7Storage(np.zeros(self.P.data.size, np.float32))       # gradWeight / gradBias views: the second flat storage
        self.o = 0

    def take(self, *size):
        n = int(np.prod(size))
        if self.o + n > self.P.data.size:
            raise T7Error("parameter vector too short for this architecture")
        w, g = T7Tensor(self.P, size, None, self.o), T7Tensor(self.G, size, None, self.o)
        self.o += n
        return w, g


def _empty():
    return T7Tensor(None, (), (), 0, "torch.FloatTensor")


def _mod(cls, **fields):
    f = {"output": _empty(), "gradInput": _empty(), "train": True, "_type": "torch.FloatTensor"}
    f.update(fields)
    return T7Object(cls, f)


def _container(cls, kids, **fields):
    return _mod(cls, modules={i + 1: k for i, k in enumerate(kids)}, **fields)


def _conv(fl, cls, ci, co, k):
    (w, gw), (b, gb) = fl.take(co, ci, k, k), fl.take(co)
    p = (k - 1) // 2
    return _mod(cls, nInputPlane=ci, nOutputPlane=co, kW=k, kH=k, dW=1, dH=1, padW=p, padH=p, weight=w, bias=b, gradWeight=gw, gradBias=gb)


def _linear(fl, i, o):
    (w, gw), (b, gb) = fl.take(o, i), fl.take(o)
    return _mod("nn.Linear", weight=w, bias=b, gradWeight=gw, gradBias=gb)


def _prelu(fl):
    w, gw = fl.take(1)
    return _mod("nn.PReLU", nOutputPlane=0, weight=w, gradWeight=gw)


def _bn(fl, cls, c, mean, var):
    (w, gw), (b, gb) = fl.take(c), fl.take(c)
    return _mod(cls, eps=1e-5, momentum=0.1, affine=True, nDim=4 if cls.startswith("nn.Spatial") else 2, weight=w, bias=b, gradWeight=gw, gradBias=gb,
                running_mean=T7Tensor.of(np.asarray(mean, np.float32)), running_var=T7Tensor.of(np.asarray(var, np.float32)))


def tree_G(kind_c, C, nz, params, running):
    """create_G_decoder_upsampling32 (models.lua:138-160) when kind_c is false, ...32c (:196-228) when true."""
    fl, r = _Flat(params), [0]

    def run(c):
        m, v = running[r[0]:r[0] + c], running[r[0] + c:r[0] + 2 * c]; r[0] += 2 * c
        return m, v

    up = lambda: _mod("nn.SpatialUpSamplingNearest", scale_factor=2)
    view = lambda *s: _mod("nn.View", size=T7Storage(np.asarray(s, np.int64)), numElements=int(np.prod(s)))
    if kind_c:
        mods = [_linear(fl, nz, 512 * 16), _prelu(fl), view(512, 4, 4)]
        stages = [(True, 512, 512, 3), (True, 512, 256, 3), (True, 256, 128, 5)]
    else:
        mods = [_linear(fl, nz, 128 * 64), view(128, 8, 8), _prelu(fl)]
        stages = [(True, 128, 256, 5), (True, 256, 128, 5)]
    for u, ci, co, k in stages:
        if u:
            mods.append(up())
        mods.append(_conv(fl, "cudnn.SpatialConvolution", ci, co, k))
        mods.append(_bn(fl, "nn.SpatialBatchNormalization", co, *run(co)))
        mods.append(_prelu(fl))
    mods += [_conv(fl, "cudnn.SpatialConvolution", 128, C, 3), _mod("nn.Sigmoid")]
    if fl.o != fl.P.data.size or r[0] != len(running):
        raise T7Error("parameter / running-statistics vectors do not match this generator")
    return _container("nn.Sequential", mods)


def _tree_stn(fl, rot, scl, trn, size, ch):
    """createSpatialTransformer (models.lua:813-905), cuda=true wiring."""
    lrelu = lambda: _mod("nn.LeakyReLU", negative_scale=0.333, negative=_empty())   # LeakyReLU.lua:5-10
    pool = lambda: _mod("nn.SpatialAveragePooling", kW=2, kH=2, dW=2, dH=2, padW=0, padH=0, ceil_mode=False, count_include_pad=True, divide=True)
    copy = lambda a, b: _mod("nn.Copy", intype=a, outtype=b, dontCast=True)
    s4 = size // 4
    nth = (1 if rot else 0) + (1 if scl else 0) + (2 if trn else 0)
    loc = _container("nn.Sequential", [pool(), _conv(fl, "nn.SpatialConvolution", ch, 16, 3), lrelu(), _conv(fl, "nn.SpatialConvolution", 16, 16, 3), lrelu(), pool(),
                                       _mod("nn.View", size=T7Storage(np.asarray([16 * s4 * s4], np.int64)), numElements=16 * s4 * s4),
                                       _linear(fl, 16 * s4 * s4, 64), lrelu(), _linear(fl, 64, nth)])
    b1 = _container("nn.Sequential", [_mod("nn.Transpose", permutations={1: {1: 3, 2: 4}, 2: {1: 2, 2: 4}}), copy("torch.CudaTensor", "torch.FloatTensor")])
    b2 = _container("nn.Sequential", [loc, _mod("nn.AffineTransformMatrixGenerator", useRotation=bool(rot), useScale=bool(scl), useTranslation=bool(trn)),
                                      _mod("nn.AffineGridGeneratorBHWD", height=size, width=size), copy("torch.CudaTensor", "torch.FloatTensor")])
    return _container("nn.Sequential", [_container("nn.ConcatTable", [b1, b2]), _mod("nn.BilinearSamplerBHWD"), copy("torch.FloatTensor", "torch.CudaTensor"),
                                        _mod("nn.Transpose", permutations={1: {1: 2, 2: 4}, 2: {1: 3, 2: 4}})])


def tree_D(C, params):
    """create_D32_st3 (models.lua:640-711), cuda=true."""
    fl = _Flat(params)
    sdrop = lambda p: _mod("nn.SpatialDropout", p=p, noise=_empty())
    copy = lambda a, b: _mod("nn.Copy", intype=a, outtype=b, dontCast=True)
    mpool = lambda: _mod("nn.SpatialMaxPooling", kW=2, kH=2, dW=2, dH=2, padW=0, padH=0, ceil_mode=False)
    mods = [copy("torch.FloatTensor", "torch.CudaTensor"), _tree_stn(fl, True, False, False, 32, C),
            _conv(fl, "nn.SpatialConvolution", C, 64, 3), _prelu(fl), _conv(fl, "nn.SpatialConvolution", 64, 64, 3), _prelu(fl),
            _mod("nn.SpatialAveragePooling", kW=2, kH=2, dW=2, dH=2, padW=0, padH=0, ceil_mode=False, count_include_pad=True, divide=True), sdrop(0.2)]
    branches = []
    for b in range(4):
        co, k1, k2 = (64, 3, 3) if b < 3 else (128, 5, 7)
        br = [_tree_stn(fl, True, True, True, 16, 64)] if b < 3 else []
        br += [_conv(fl, "nn.SpatialConvolution", 64, co, k1), _prelu(fl), mpool(), sdrop(0.2), _conv(fl, "nn.SpatialConvolution", co, co, k2), _prelu(fl)]
        branches.append(_container("nn.Sequential", br))
    mods += [_container("nn.Concat", branches, dimension=2, size=T7Storage(np.zeros(0, np.int64))), sdrop(0.5),
             _mod("nn.View", size=T7Storage(np.asarray([20480], np.int64)), numElements=20480), _linear(fl, 20480, 256), _prelu(fl),
             _mod("nn.Dropout", p=0.5, v2=True, inplace=False, noise=_empty()), _linear(fl, 256, 1), _mod("nn.Sigmoid"), copy("torch.CudaTensor", "torch.FloatTensor")]
    if fl.o != fl.P.data.size:
        raise T7Error("parameter vector does not match create_D32_st3")
    return _container("nn.Sequential", mods)


def tree_V(C, params, running):
    """create_V32 (models.lua:765-804)."""
    fl, r = _Flat(params), [0]

    def run(c):
        m, v = running[r[0]:r[0] + c], running[r[0] + c:r[0] + 2 * c]; r[0] += 2 * c
        return m, v

    lrelu = lambda: _mod("nn.LeakyReLU", negative_scale=0.333, negative=_empty())   # LeakyReLU.lua:5-10
    mpool = lambda: _mod("nn.SpatialMaxPooling", kW=2, kH=2, dW=2, dH=2, padW=0, padH=0, ceil_mode=False)
    drop = lambda: _mod("nn.Dropout", p=0.5, v2=True, inplace=False, noise=_empty())
    mods = [_conv(fl, "nn.SpatialConvolution", C, 128, 3), lrelu(), mpool(),
            _conv(fl, "nn.SpatialConvolution", 128, 128, 3), _bn(fl, "nn.SpatialBatchNormalization", 128, *run(128)), lrelu(), mpool(), drop(),
            _conv(fl, "nn.SpatialConvolution", 128, 256, 3), lrelu(),
            _conv(fl, "nn.SpatialConvolution", 256, 256, 3), _bn(fl, "nn.SpatialBatchNormalization", 256, *run(256)), lrelu(), mpool(),
            _mod("nn.SpatialDropout", p=0.5, noise=_empty()), _mod("nn.View", size=T7Storage(np.asarray([4096], np.int64)), numElements=4096),
            _linear(fl, 4096, 1024), _bn(fl, "nn.BatchNormalization", 1024, *run(1024)), lrelu(), drop(),
            _linear(fl, 1024, 1024), _bn(fl, "nn.BatchNormalization", 1024, *run(1024)), lrelu(), drop(),
            _linear(fl, 1024, 2), _mod("nn.SoftMax")]
    if fl.o != fl.P.data.size or r[0] != len(running):
        raise T7Error("parameter / running-statistics vectors do not match create_V32")
    return _container("nn.Sequential", mods)
