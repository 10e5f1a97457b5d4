"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  PARITY UNPINNED: see oracle/catgen_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcatgen_oracle.so")
_SO64 = os.path.join(_HERE, "libcatgen_oracle_f64.so")   # the same source with float widened to double (catgen_oracle.h, OG_F64)

G32UP, G32UPC, D32_ST3 = 0, 1, 2
V32 = 3   # create_V32 (models.lua:765-804); restated below as a composition of the og_* operators, forward / evaluate() only
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("catgen_oracle.c", "catgen_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or not os.path.exists(_SO64) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class StepCfg(C.Structure):
    _fields_ = [("B", C.c_int), ("d_iters", C.c_int), ("g_iters", C.c_int),
                ("D_L1", C.c_float), ("D_L2", C.c_float), ("G_L1", C.c_float), ("G_L2", C.c_float),
                ("D_clamp", C.c_float), ("G_clamp", C.c_float),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


def default_cfg(B, d_iters=1, g_iters=1):
    """train.lua:20-48 defaults: D_L1=0 D_L2=1e-4 G_L1=G_L2=0 D_clamp=1 G_clamp=5; optim.adam defaults."""
    return StepCfg(B, d_iters, g_iters, 0.0, 1e-4, 0.0, 0.0, 1.0, 5.0, 1e-3, 0.9, 0.999, 1e-8)


def usable_cpus():
    """CPUs this process may actually use: affinity mask, further limited by a cgroup v2 CPU quota if one is set.
    os.cpu_count() reports the machine (128 on the B200 host) and oversubscribing made the oracle crawl."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per))))
    except Exception:
        pass
    return n


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")   # idle workers sleep instead of spinning between the many short regions
    L = C.CDLL(_SO)
    i, l, f, vp = C.c_int, C.c_long, C.c_float, C.c_void_p

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("og_set_threads", None, i)
    sig("og_get_threads", i)
    sig("og_conv2d_fwd", None, _fp, _fp, _fp, _fp, i, i, i, i, i, i)
    sig("og_conv2d_bwd_data", None, _fp, _fp, _fp, i, i, i, i, i, i)
    sig("og_conv2d_bwd_filter", None, _fp, _fp, _fp, _fp, i, i, i, i, i, i)
    sig("og_linear_fwd", None, _fp, _fp, _fp, _fp, i, i, i)
    sig("og_linear_bwd", None, _fp, _fp, _fp, _fp, _fp, _fp, i, i, i)
    sig("og_bn_fwd_train", None, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, i, i, i, f, f)
    sig("og_bn_fwd_eval", None, _fp, _fp, _fp, _fp, _fp, _fp, i, i, i, f)
    sig("og_bn_bwd_train", None, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, i, i, i)
    sig("og_prelu_fwd", None, _fp, f, _fp, l)
    sig("og_prelu_bwd", None, _fp, _fp, f, _fp, _fp, l)
    sig("og_leakyrelu_fwd", None, _fp, f, _fp, l)
    sig("og_leakyrelu_bwd", None, _fp, _fp, f, _fp, l)
    sig("og_upsample2x_fwd", None, _fp, _fp, i, i, i)
    sig("og_upsample2x_bwd", None, _fp, _fp, i, i, i)
    sig("og_sigmoid_fwd", None, _fp, _fp, l)
    sig("og_sigmoid_bwd", None, _fp, _fp, _fp, l)
    sig("og_avgpool2_fwd", None, _fp, _fp, i, i, i)
    sig("og_avgpool2_bwd", None, _fp, _fp, i, i, i)
    sig("og_maxpool2_fwd", None, _fp, _fp, _ip, i, i, i)
    sig("og_maxpool2_bwd", None, _fp, _ip, _fp, i, i, i)
    sig("og_mask_channels", None, _fp, _fp, _fp, i, i)
    sig("og_affine_matrix_fwd", None, _fp, _fp, i, i, i, i)
    sig("og_affine_matrix_bwd", None, _fp, _fp, _fp, i, i, i, i)
    sig("og_affine_grid_fwd", None, _fp, _fp, i, i, i)
    sig("og_affine_grid_bwd", None, _fp, _fp, i, i, i)
    sig("og_bilinear_fwd", None, _fp, _fp, _fp, i, i, i, i)
    sig("og_bilinear_bwd", None, _fp, _fp, _fp, _fp, _fp, i, i, i, i)
    sig("og_bce_fwd", f, _fp, _fp, i)
    sig("og_bce_bwd", None, _fp, _fp, _fp, i)
    sig("og_adam_step", None, _fp, _fp, _fp, _fp, l, i, f, f, f, f)
    sig("og_conv_upsample_fwd", None, _fp, _fp, _fp, _fp, i, i, i, i, i, i, i)
    sig("og_model_create", vp, i, i, i)
    sig("og_model_free", None, vp)
    sig("og_model_nparams", l, vp)
    sig("og_model_params", _fp, vp)
    sig("og_model_grads", _fp, vp)
    sig("og_model_bn_running", _fp, vp, C.POINTER(l))
    sig("og_model_init", None, vp, C.c_ulonglong)
    sig("og_model_zero_grads", None, vp)
    sig("og_D_mask_floats", l, i)
    sig("og_G_forward", None, vp, _fp, i, _fp, i)
    sig("og_G_backward", None, vp, _fp, _fp)
    sig("og_D_forward", None, vp, _fp, i, _fp, _fp, _fp)
    sig("og_D_backward", None, vp, _fp, _fp)
    sig("og_trainer_create", vp, vp, vp)
    sig("og_trainer_free", None, vp)
    sig("og_train_step", None, vp, C.POINTER(StepCfg), _fp, _fp, _fp, _fp, _fp, _fp, _fp)
    sig("og_fevalD", f, vp, C.POINTER(StepCfg), _fp, _fp, _fp, _fp)
    sig("og_fevalG_on_D", f, vp, C.POINTER(StepCfg), _fp, _fp)
    L.og_set_threads(min(usable_cpus(), 16))   # tests use batches <= 8; bench.py raises this explicitly
    _lib = L
    return L


_lib64 = None


def lib64():
    """The float64 build: model-level entry points only (og_step_cfg has another layout there and the trainer is not bound)."""
    global _lib64
    if _lib64 is not None:
        return _lib64
    if not os.path.exists(_SO64):
        build(force=True)
    L = C.CDLL(_SO64)
    i, l, vp, dp = C.c_int, C.c_long, C.c_void_p, C.POINTER(C.c_double)

    def sig(name, res, *args):
        fn = getattr(L, name); fn.restype = res; fn.argtypes = list(args)
    sig("og_model_create", vp, i, i, i); sig("og_model_free", None, vp); sig("og_model_nparams", l, vp)
    sig("og_model_params", dp, vp); sig("og_model_grads", dp, vp); sig("og_model_zero_grads", None, vp)
    sig("og_D_mask_floats", l, i)
    sig("og_G_forward", None, vp, dp, i, dp, i); sig("og_G_backward", None, vp, dp, dp)
    sig("og_D_forward", None, vp, dp, i, dp, dp, dp); sig("og_D_backward", None, vp, dp, dp)
    L.og_set_threads(min(usable_cpus(), 16))
    _lib64 = L
    return L


def P(a):
    """float32 C-contiguous ndarray (or None) -> float*"""
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_fp)


def IP(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_ip)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _OwnedView(np.ndarray):
    """ndarray view into memory owned by an oracle Model; keeps that Model alive (a bare view of a temporary
    `po.Model(...).params` dangled once the temporary was collected -- found as +inf losses in a GPU test)."""
    _owner = None


def _view(ptr, n, owner):
    v = np.ctypeslib.as_array(ptr, shape=(n,)).view(_OwnedView)
    v._owner = owner
    return v


class Model:
    """Oracle G or D with parameters as a flat vector in nn getParameters() order.
    f64=True: the float64 build (parameters start at zero there -- copy them from an fp32 Model; forward/backward only)."""

    def __init__(self, kind, C_img=3, nz=100, seed=None, f64=False):
        self.f64 = bool(f64)
        self.L = lib64() if f64 else lib()
        self.dt = np.float64 if f64 else np.float32
        assert not (f64 and seed is not None), "initialise an fp32 Model and copy its parameters"
        self.kind, self.C, self.nz = kind, C_img, nz
        self.h = self.L.og_model_create(kind, C_img, nz)
        self.n = self.L.og_model_nparams(self.h)
        if seed is not None:
            self.L.og_model_init(self.h, seed)

    def __del__(self):
        try:
            self.L.og_model_free(self.h)
        except Exception:
            pass

    @property
    def params(self):
        return _view(self.L.og_model_params(self.h), self.n, self)

    @property
    def grads(self):
        return _view(self.L.og_model_grads(self.h), self.n, self)

    @property
    def bn_running(self):
        assert not self.f64
        n = C.c_long()
        p = self.L.og_model_bn_running(self.h, C.byref(n))
        return _view(p, n.value, self) if n.value else np.zeros(0, np.float32)

    def zero_grads(self):
        self.L.og_model_zero_grads(self.h)

    def _a(self, a):
        return np.ascontiguousarray(a, dtype=self.dt)

    def _p(self, a):
        if a is None:
            return None
        assert a.dtype == self.dt and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.POINTER(C.c_double if self.f64 else C.c_float))

    def G_forward(self, z, train=True):
        z = self._a(z)
        B = z.shape[0]
        out = np.empty((B, self.C, 32, 32), self.dt)
        self.L.og_G_forward(self.h, self._p(z), B, self._p(out), 1 if train else 0)
        return out

    def G_backward(self, gout):
        gout = self._a(gout)
        gz = np.empty((gout.shape[0], self.nz), self.dt)
        self.L.og_G_backward(self.h, self._p(gout), self._p(gz))
        return gz

    def D_forward(self, x, masks=None):
        x = self._a(x)
        B = x.shape[0]
        sig, pre = np.empty(B, self.dt), np.empty(B, self.dt)
        if masks is not None:
            masks = self._a(masks)
            assert masks.size == self.L.og_D_mask_floats(B)
        self.L.og_D_forward(self.h, self._p(x), B, self._p(masks), self._p(sig), self._p(pre))
        return sig, pre

    def D_backward(self, gout):
        gout = self._a(gout)
        gx = np.empty((gout.shape[0], self.C, 32, 32), self.dt)
        self.L.og_D_backward(self.h, self._p(gout), self._p(gx))
        return gx


V32_BN = (128, 256, 1024, 1024)


def V32_nparams(C_img):
    """getParameters() length of create_V32 (models.lua:765-804): 4 convs, 3 Linear, 4 BatchNormalization (gamma, beta)."""
    n = 0
    for ci, co in ((C_img, 128), (128, 128), (128, 256), (256, 256)):
        n += co * ci * 9 + co
    n += 1024 * 4096 + 1024 + 1024 * 1024 + 1024 + 2 * 1024 + 2
    return n + 2 * sum(V32_BN)


def V32_init(C_img, seed):
    """weight-init 'heuristic' shape on V: W ~ U(+-1/sqrt(fan_in)) (placeholder distribution for tests: V is loaded from
    a trained v_*.net in the reference, train.lua:119-123, never used at its initial values), bias 0, gamma ~ U(0,1), beta 0;
    running mean ~ N(0, 0.02), running var ~ U(0.04, 0.12) so that evaluate()-mode BN is exercised with non-trivial statistics and the
    outputs differ visibly between images."""
    rng = np.random.default_rng(seed)
    parts, run = [], []

    def wb(co, fan_in):
        s = 1.0 / np.sqrt(fan_in)
        parts.append(rng.uniform(-s, s, co * fan_in).astype(np.float32))
        parts.append(rng.uniform(-0.05, 0.05, co).astype(np.float32))

    def bn(c):
        parts.append(rng.uniform(0.5, 1.5, c).astype(np.float32))
        parts.append(rng.uniform(-0.1, 0.1, c).astype(np.float32))
        run.append(rng.normal(0, 0.02, c).astype(np.float32))
        run.append(rng.uniform(0.04, 0.12, c).astype(np.float32))   # about the variance the layer's input really has at these weights: keeps the signal alive through seven layers

    wb(128, C_img * 9); wb(128, 128 * 9); bn(128); wb(256, 128 * 9); wb(256, 256 * 9); bn(256)
    wb(1024, 4096); bn(1024); wb(1024, 1024); bn(1024); wb(2, 1024)
    flat = np.concatenate(parts)
    assert flat.size == V32_nparams(C_img)
    return flat, np.concatenate(run)


def V_forward(flat, running, x):
    """MODEL_V:forward(images) in evaluate() mode (train.lua:123; utils/nn_utils.lua:700), NCHW, through the oracle's
    operators: conv-LeakyReLU(0.333)-maxpool, conv-BN-LReLU-maxpool-Dropout(id), conv-LReLU, conv-BN-LReLU-maxpool-
    SpatialDropout(x0.5)-View, Linear-BN-LReLU-Dropout(id) x2, Linear-SoftMax (models.lua:769-799).  Returns [B,2]."""
    L = lib()
    x = np.ascontiguousarray(x, np.float32)
    B, Ci = x.shape[0], x.shape[1]
    o, r = [0], [0]

    def take(n):
        a = np.ascontiguousarray(flat[o[0]:o[0] + n], np.float32); o[0] += n
        return a

    def run(n):
        a = np.ascontiguousarray(running[r[0]:r[0] + n], np.float32); r[0] += n
        return a

    def conv(h, co):
        n, ci, hh, ww = h.shape
        W, b = take(co * ci * 9), take(co)
        y = np.empty((n, co, hh, ww), np.float32)
        L.og_conv2d_fwd(P(h), P(W), P(b), P(y), n, ci, hh, ww, co, 3)
        return y

    def bn(h):
        n, c = h.shape[0], h.shape[1]
        hw = int(np.prod(h.shape[2:])) if h.ndim > 2 else 1
        g, b, m, v = take(c), take(c), run(c), run(c)
        y = np.empty_like(h)
        L.og_bn_fwd_eval(P(h), P(g), P(b), P(y), P(m), P(v), n, c, hw, 1e-5)
        return y

    def lrelu(h):
        y = np.empty_like(h)
        L.og_leakyrelu_fwd(P(h), 0.333, P(y), h.size)   # LeakyReLU.lua:5-10 (row A8), as in D's localisation networks
        return y

    def pool(h):
        n, c, hh, ww = h.shape
        y = np.empty((n, c, hh // 2, ww // 2), np.float32)
        idx = np.empty(y.shape, np.int32)
        L.og_maxpool2_fwd(P(h), P(y), IP(idx), n * c, hh, ww)
        return y

    def linear(h, out):
        n, k = h.shape
        W, b = take(out * k), take(out)
        y = np.empty((n, out), np.float32)
        L.og_linear_fwd(P(h), P(W), P(b), P(y), n, k, out)
        return y

    h = pool(lrelu(conv(x, 128)))
    h = pool(lrelu(bn(conv(h, 128))))
    h = lrelu(conv(h, 256))
    h = pool(lrelu(bn(conv(h, 256)))) * np.float32(0.5)
    h = np.ascontiguousarray(h.reshape(B, 4096))
    h = lrelu(bn(linear(h, 1024)))
    h = lrelu(bn(linear(h, 1024)))
    h = linear(h, 2)
    assert o[0] == flat.size and r[0] == running.size
    e = np.exp(h - h.max(axis=1, keepdims=True))      # nn.SoftMax: shift by the row max, exp, normalise
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def rateWithV(flat, running, images):
    """utils/nn_utils.lua:686-711: 1 - mean(predictions[i][1]) (first neuron = P(fake))."""
    return 1.0 - float(V_forward(flat, running, images)[:, 0].astype(np.float64).mean())


def D_mask_floats(B):
    return lib().og_D_mask_floats(B)


def make_D_masks(B, rng):
    """Train-mode dropout multipliers in the layout og_D_forward expects."""
    sp = (rng.random(B * (64 * 4 + 128)) >= 0.2).astype(np.float32)          # 5x SpatialDropout(0.2), no rescale
    head = (rng.random(B * 320) >= 0.5).astype(np.float32)                   # SpatialDropout(0.5)
    fc = (rng.random(B * 256) >= 0.5).astype(np.float32) * 2.0               # Dropout(0.5), rescaled by 1/(1-p)
    return np.concatenate([sp, head, fc]).astype(np.float32)


class Trainer:
    def __init__(self, G, D):
        self.L = lib()
        self.G, self.D = G, D
        self.h = self.L.og_trainer_create(G.h, D.h)

    def __del__(self):
        try:
            self.L.og_trainer_free(self.h)
        except Exception:
            pass

    def step(self, cfg, real, zD, zG, masks=None):
        real, zD, zG = f32(real), f32(zD), f32(zG)
        lossD = np.zeros(cfg.d_iters, np.float32)
        lossG = np.zeros(cfg.g_iters, np.float32)
        d_out = np.zeros(cfg.B, np.float32)
        if masks is not None:
            masks = f32(masks)
        self.L.og_train_step(self.h, C.byref(cfg), P(real), P(zD), P(zG), P(masks), P(lossD), P(lossG), P(d_out))
        return lossD, lossG, d_out
