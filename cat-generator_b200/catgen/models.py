"""Mirror of the reference's models.lua surface (models.create_G / models.create_D) over libcatgen.

Reference: /root/reference/models.lua:234-240 (create_G), :268-277 (create_D).  The returned objects follow the
Torch7 nn.Module protocol the reference's callers use (train.lua:147-185, adversarial.lua:84-89,187-197):
forward / backward / getParameters / training / evaluate.  All compute happens in libcatgen's CUDA kernels.
"""
import ctypes as C

import numpy as np

from . import lib
from .lib import check, P, f32


class Module:
    """One G or D network living on the GPU behind a cg_model handle."""

    def __init__(self, kind, C_img, nz=100, seed=1):
        self.L = lib.load()
        self.kind, self.C, self.nz = kind, int(C_img), int(nz)
        self.h = C.c_void_p()
        check(self.L.cg_model_create(C.byref(self.h), kind, self.C, self.nz, int(seed)))
        n = C.c_int64()
        check(self.L.cg_model_nparams(self.h, C.byref(n)))
        self.nparams = n.value
        self.output = None
        self.gradInput = None
        self._B = 0

    def __del__(self):
        try:
            self.L.cg_model_free(self.h)
        except Exception:
            pass

    # -- MODEL:getParameters() (train.lua:184-185): flat copies, nn getParameters() order
    def getParameters(self):
        return self.get_params(), self.get_grads()

    def get_params(self):
        a = np.empty(self.nparams, np.float32)
        check(self.L.cg_model_get_params(self.h, P(a)))
        return a

    def set_params(self, a):
        a = f32(a)
        assert a.size == self.nparams
        check(self.L.cg_model_set_params(self.h, P(a)))

    def get_grads(self):
        a = np.empty(self.nparams, np.float32)
        check(self.L.cg_model_get_grads(self.h, P(a)))
        return a

    def zeroGradParameters(self):
        check(self.L.cg_model_zero_grads(self.h))

    def training(self):
        check(self.L.cg_model_set_mode(self.h, 1))

    def evaluate(self):
        check(self.L.cg_model_set_mode(self.h, 0))

    def get_bn_running(self):
        n = C.c_int64()
        check(self.L.cg_model_bn_running_len(self.h, C.byref(n)))
        a = np.empty(n.value, np.float32)
        if n.value:
            check(self.L.cg_model_get_bn_running(self.h, P(a)))
        return a

    def set_bn_running(self, a):
        check(self.L.cg_model_set_bn_running(self.h, P(f32(a))))


class Generator(Module):
    def forward(self, z):
        z = f32(z)
        B = z.shape[0]
        out = np.empty((B, self.C, 32, 32), np.float32)
        check(self.L.cg_G_forward(self.h, P(z), B, P(out)))
        self._B, self.output = B, out
        return out

    def backward(self, z, gradOutput):
        g = f32(gradOutput)
        gz = np.empty((self._B, self.nz), np.float32)
        check(self.L.cg_G_backward(self.h, P(g), P(gz)))
        self.gradInput = gz
        return gz


class Discriminator(Module):
    def forward(self, x, with_pre=False):
        x = f32(x)
        B = x.shape[0]
        sig, pre = np.empty(B, np.float32), np.empty(B, np.float32)
        check(self.L.cg_D_forward(self.h, P(x), B, P(sig), P(pre)))
        self._B, self.output, self.pre_sigmoid = B, sig.reshape(B, 1), pre
        return (self.output, pre) if with_pre else self.output

    def backward(self, x, gradOutput):
        g = f32(np.asarray(gradOutput).reshape(-1))
        gx = np.empty((self._B, self.C, 32, 32), np.float32)
        check(self.L.cg_D_backward(self.h, P(g), P(gx)))
        self.gradInput = gx          # adversarial.lua:193 reads MODEL_D.modules[1].gradInput
        return gx

    def get_masks(self):
        n = C.c_int64()
        check(self.L.cg_D_mask_floats(self._B, C.byref(n)))
        a = np.empty(n.value, np.float32)
        check(self.L.cg_D_get_masks(self.h, P(a)))
        return a

    def set_masks(self, masks, B, count=1):
        check(self.L.cg_D_set_masks(self.h, P(f32(masks)), int(B), int(count)))


class Validator(Module):
    """create_V32 (models.lua:765-804).  Forward only and always in evaluate() mode, the way train.lua:119-123 uses it."""

    def forward(self, x):
        x = f32(x)
        B = x.shape[0]
        out = np.empty((B, 2), np.float32)
        check(self.L.cg_V_forward(self.h, P(x), B, P(out)))
        self._B, self.output = B, out
        return out

    def training(self):
        raise lib.CatgenError("V is only ever evaluated on this path (train.lua:123); train_v.lua is out of scope")


def create_V(dimensions, seed=3):
    """models.create_V(dimensions), models.lua:716-721 -> create_V32 at 32x32."""
    if dimensions[1] != 32 or dimensions[2] != 32:
        raise lib.CatgenError("only create_V32 is built (SURVEY.md section 8 row F3)")
    return Validator(lib.V32, dimensions[0], 100, seed)


def create_G(dimensions, noiseDim, seed=1, kind=None):
    """models.create_G(dimensions, noiseDim), models.lua:234-240.  32x32 -> G32up-c; `kind` selects G32up."""
    if dimensions[1] != 32 or dimensions[2] != 32:
        raise lib.CatgenError("only the 32x32 generators are on the hot path (SURVEY.md section 2, row 15)")
    return Generator(lib.G32UPC if kind is None else kind, dimensions[0], noiseDim, seed)


def create_D(dimensions, cuda=True, seed=2):
    """models.create_D(dimensions, cuda), models.lua:268-277 -> create_D32_st3.  `cuda` must be truthy."""
    if dimensions[1] != 32 or dimensions[2] != 32:
        raise lib.CatgenError("only create_D32_st3 is on the hot path (SURVEY.md section 2, row 15)")
    if not cuda:
        raise lib.CatgenError("libcatgen has no CPU path; create_D(dimensions, cuda=false) is not available")
    return Discriminator(lib.D32_ST3, dimensions[0], 100, seed)
