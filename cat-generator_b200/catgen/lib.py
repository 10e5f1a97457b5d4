"""ctypes binding of libcatgen.so (include/catgen.h).  This is the Python twin of lua/catgen_ffi.lua.

The product path has NO CPU fallback: if the shared library is missing, or no sm_100 CUDA device is
present, loading / cg_init raises -- nothing here ever routes to the oracle or to PyTorch ops.
"""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(_HERE)
SO_PATH = os.path.join(PKG_DIR, "libcatgen.so")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "catgen.h")

G32UP, G32UPC, D32_ST3 = 0, 1, 2
V32 = 3
fp = C.POINTER(C.c_float)


class CatgenError(RuntimeError):
    pass


class StepCfg(C.Structure):
    """cg_step_cfg; defaults are train.lua:26-36 and optim.adam's."""
    _fields_ = [("B", C.c_int), ("d_iters", C.c_int), ("g_iters", C.c_int),
                ("D_L1", C.c_float), ("D_L2", C.c_float), ("G_L1", C.c_float), ("G_L2", C.c_float),
                ("D_clamp", C.c_float), ("G_clamp", C.c_float),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


def default_cfg(B, d_iters=1, g_iters=1):
    return StepCfg(B, d_iters, g_iters, 0.0, 1e-4, 0.0, 0.0, 1.0, 5.0, 1e-3, 0.9, 0.999, 1e-8)


def declared_symbols(header=HEADER):
    """Every function name include/catgen.h declares (used by the symbol-export test)."""
    src = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", src)))


_lib = None


def load():
    """dlopen libcatgen.so.  Raises CatgenError if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise CatgenError("%s not found: build it with `make -C %s` (or __graft_entry__.build()); "
                          "there is no CPU fallback" % (SO_PATH, PKG_DIR))
    L = C.CDLL(SO_PATH)
    L.cg_last_error.restype = C.c_char_p
    L.cg_version.restype = C.c_char_p
    L.cg_launch_count.restype = C.c_int64
    L.cg_reset_launch_count.restype = None
    L.cg_shutdown.restype = None
    L.cg_dev_alloc.restype = C.c_void_p
    L.cg_dev_alloc.argtypes = [C.c_int64]
    L.cg_dev_free.argtypes = [C.c_void_p]
    L.cg_host_alloc.restype = C.c_void_p
    L.cg_host_alloc.argtypes = [C.c_int64]
    L.cg_host_free.argtypes = [C.c_void_p]
    L.cg_dev_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.cg_dev_download.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.cg_uniform_dev.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_uint64, C.c_uint64]
    L.cg_model_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_uint64]
    L.cg_trainer_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
    L.cg_train_step_dev.argtypes = [C.c_void_p, C.POINTER(StepCfg), C.c_void_p, C.c_void_p, C.c_void_p, fp, fp]
    L.cg_train_step.argtypes = [C.c_void_p, C.POINTER(StepCfg), C.c_void_p, C.c_void_p, C.c_void_p, fp, fp, fp]
    L.cg_penalty_clamp.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, fp]
    L.cg_set_precision.argtypes = [C.c_int]
    L.cg_prelu_fwd.argtypes = [fp, C.c_float, fp, C.c_int64]
    L.cg_prelu_bwd.argtypes = [fp, fp, C.c_float, fp, fp, C.c_int64]
    L.cg_leakyrelu_fwd.argtypes = [fp, C.c_float, fp, C.c_int64]
    L.cg_leakyrelu_bwd.argtypes = [fp, fp, C.c_float, fp, C.c_int64]
    L.cg_sigmoid_fwd.argtypes = [fp, fp, C.c_int64]
    L.cg_sigmoid_bwd.argtypes = [fp, fp, fp, C.c_int64]
    _lib = L
    return L


def check(status):
    if status != 0:
        raise CatgenError("libcatgen error %d: %s" % (status, load().cg_last_error().decode()))


def init(device=0):
    """cg_init; raises (loudly) when there is no sm_100 device -- the library has no CPU path."""
    check(load().cg_init(int(device)))


def P(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "need C-contiguous float32"
    return a.ctypes.data_as(fp)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)
