"""catgen -- Python mirror of the reference's Lua module surface over libcatgen.so (sm_100a CUDA kernels).

models.create_G / models.create_D, the nn.Module-style forward/backward protocol and adversarial.train are
re-expressed in catgen.models / catgen.nn / catgen.adversarial on top of the C-ABI in include/catgen.h.
"""
from . import lib  # noqa: F401
from .lib import CatgenError, StepCfg, default_cfg, G32UP, G32UPC, D32_ST3, V32  # noqa: F401
