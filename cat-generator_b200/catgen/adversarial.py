"""Mirror of adversarial.train (/root/reference/adversarial.lua:27-292) driving libcatgen's fused step.

The per-sample Lua loops of the reference (:101-106, :225-238) are batched into one cg_train_step call per
step; the order of operations inside the step is the reference's (see csrc/capi.cu train_step_core).
"""
import ctypes as C
import time

import numpy as np

from . import lib
from .lib import check, P, f32


class Trainer:
    """Holds the Adam state of one (G, D) pair: OPTSTATE.adam.D / .G (train.lua:191-207)."""

    def __init__(self, G, D):
        self.L = lib.load()
        self.G, self.D = G, D
        self.h = C.c_void_p()
        check(self.L.cg_trainer_create(C.byref(self.h), G.h, D.h))

    def __del__(self):
        try:
            self.L.cg_trainer_free(self.h)
        except Exception:
            pass

    def step(self, cfg, real, zD, zG):
        real, zD, zG = f32(real), f32(zD), f32(zG)
        lossD = np.zeros(cfg.d_iters, np.float32)
        lossG = np.zeros(cfg.g_iters, np.float32)
        d_out = np.zeros(cfg.B, np.float32)
        check(self.L.cg_train_step(self.h, C.byref(cfg), real.ctypes.data, zD.ctypes.data, zG.ctypes.data,
                                   P(lossD), P(lossG), P(d_out)))
        return lossD, lossG, d_out


class Opt:
    """The OPT fields adversarial.train reads, with train.lua:17-48 defaults."""

    def __init__(self, **kw):
        self.batchSize, self.N_epoch, self.noiseDim = 32, 1000, 100
        self.D_L1, self.D_L2, self.G_L1, self.G_L2 = 0.0, 1e-4, 0.0, 0.0
        self.D_clamp, self.G_clamp = 1.0, 5.0
        self.D_iterations, self.G_iterations = 1, 1
        self.seed = 1
        self.__dict__.update(kw)


def train(trainer, opt, trainData, rng, maxAccuracyD=1.01, accsInterval=20, log=None):
    """One epoch (adversarial.lua:27-292).  trainData: float32 [n,C,32,32] in [0,1].  Returns
    (totalValid, seconds): D's accuracy over the epoch's confusion matrix and the wall time the reference
    prints (:278-280).  maxAccuracyD <= 1 (D-skipping, :157-166) is not supported: with stock optim.adam the
    reference itself cannot take that branch (feval returning false), and the default is 1.01."""
    if maxAccuracyD <= 1.0:
        raise lib.CatgenError("D_maxAcc <= 1 is not supported (see docstring)")
    N_epoch = opt.N_epoch if opt.N_epoch > 0 else trainData.shape[0]
    B, half = opt.batchSize, opt.batchSize // 2
    confusion = np.zeros((2, 2), np.int64)
    t0 = time.time()
    for t in range(1, N_epoch + 1, half):                                    # :51
        thisB = min(B, N_epoch - t + 1)                                      # :53
        if thisB < 4:                                                        # :65-68
            break
        thisB -= thisB % 2
        cfg = lib.StepCfg(thisB, opt.D_iterations, opt.G_iterations, opt.D_L1, opt.D_L2, opt.G_L1, opt.G_L2,
                          opt.D_clamp, opt.G_clamp, 1e-3, 0.9, 0.999, 1e-8)
        idx = rng.integers(0, trainData.shape[0], (opt.D_iterations, thisB // 2))          # :225-230
        real = trainData[idx]
        zD = rng.uniform(-1, 1, (opt.D_iterations, thisB // 2, opt.noiseDim)).astype(np.float32)   # :233
        zG = rng.uniform(-1, 1, (opt.G_iterations, thisB, opt.noiseDim)).astype(np.float32)        # :254
        _, _, d_out = trainer.step(cfg, real, zD, zG)
        pred = (d_out > 0.5).astype(np.int64)                                # :101-106
        tgt = np.concatenate([np.ones(thisB // 2, np.int64), np.zeros(thisB - thisB // 2, np.int64)])
        np.add.at(confusion, (pred, tgt), 1)
    dt = time.time() - t0
    if log:
        log("<trainer> time required for this epoch = %d s" % dt)
        log("<trainer> time to learn 1 sample = %f ms" % (1000 * dt / N_epoch))
    total = confusion.sum()
    return (np.trace(confusion) / total if total else 0.0), dt
