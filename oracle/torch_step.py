"""PyTorch-CPU fp32 restatement of ONE adversarial.train loop body (TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE).

BASELINE.md section 3 "B-mkl": the reference's Torch7 `nn` CPU path cannot run here (no Lua/LuaJIT/Torch7), and the C oracle's
hand-rolled SGEMM is far below what Torch7 linked against an optimised BLAS reaches.  This module states the same step with
PyTorch CPU ops (oneDNN/MKL) + autograd as the stand-in for "Torch7 nn + optimised BLAS" so that the CPU figure printed next to
the GPU one is a fair one.  Only bench.py's cpu_baseline / --impl reference legs and tests/ may import it.  PARITY UNPINNED
(oracle/catgen_oracle.h).

Step order follows /root/reference/adversarial.lua:221-266 (fevalD :72-167, fevalG_on_D :171-215, optim.adam [upstream],
SURVEY.md A.8): D sees B/2 real then B/2 fake images produced by a SEPARATE G forward; penalties before the clamp; Adam with eps
added before the bias correction.  In fevalG the reference's MODEL_D:backward also accumulates D's (unused) parameter gradients;
they are computed here too, so the timed work is the reference's.
"""
import math

import torch

from . import torch_ref as tr


class Net:
    def __init__(self, flat):
        self.p = torch.as_tensor(flat, dtype=torch.float32).clone().requires_grad_(True)
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.t = 0


def _penalty_clamp_adam(net, g, l1, l1sign, l2, clampv, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    with torch.no_grad():
        p = net.p
        add = 0.0
        if l1 != 0 or l2 != 0:                                   # adversarial.lua:92-98 / :201-208
            add = float(l1 * p.abs().sum() + l2 * (p * p).sum() / 2)
            g = g + torch.sign(p) * l1sign + p * l2
        if clampv != 0:
            g = g.clamp(-clampv, clampv)                         # :110-112 / :210-212
        net.t += 1
        net.m.mul_(b1).add_(g, alpha=1 - b1)
        net.v.mul_(b2).addcmul_(g, g, value=1 - b2)
        step = lr * math.sqrt(1 - b2 ** net.t) / (1 - b1 ** net.t)
        p.addcdiv_(net.m, net.v.sqrt().add_(eps), value=-step)  # eps BEFORE the bias correction (optim.adam)
    return add, g


def train_step(G, D, cfg, real, zD, zG, masks=None, kind="G32UPC", C=3):
    """real [d][B/2,C,32,32], zD [d][B/2,nz], zG [g][B,nz], masks [(d+g)][n] or None (evaluate()-mode dropout).
    Returns lossD[d], lossG[g], d_out[B] (D's outputs of the last D update)."""
    B, hB = cfg.B, cfg.B // 2
    lossD, lossG, d_out, mi = [], [], None, 0
    as_t = lambda a: torch.as_tensor(a, dtype=torch.float32)
    for k in range(cfg.d_iters):
        with torch.no_grad():
            fake = tr.G_forward(G.p, as_t(zD[k]), kind, C, zD.shape[-1])
        x = torch.cat([as_t(real[k]), fake], 0)
        tgt = torch.cat([torch.ones(hB), torch.zeros(B - hB)])
        out, _ = tr.D_forward(D.p, x, None if masks is None else as_t(masks[mi]), C); mi += 1
        loss = tr.bce(out, tgt)
        (g,) = torch.autograd.grad(loss, [D.p])
        add, _ = _penalty_clamp_adam(D, g, cfg.D_L1, cfg.D_L1, cfg.D_L2, cfg.D_clamp, cfg.lr, cfg.beta1, cfg.beta2, cfg.eps)
        lossD.append(float(loss.detach()) + add); d_out = out.detach().numpy().copy()
    for k in range(cfg.g_iters):
        samples = tr.G_forward(G.p, as_t(zG[k]), kind, C, zG.shape[-1])
        out, _ = tr.D_forward(D.p, samples, None if masks is None else as_t(masks[mi]), C); mi += 1
        loss = tr.bce(out, torch.ones(B))
        g, _dead = torch.autograd.grad(loss, [G.p, D.p])        # _dead: D's gradients, computed and dropped like the reference
        add, _ = _penalty_clamp_adam(G, g, cfg.G_L1, cfg.G_L2, cfg.G_L2, cfg.G_clamp, cfg.lr, cfg.beta1, cfg.beta2, cfg.eps)
        lossG.append(float(loss.detach()) + add)
    return lossD, lossG, d_out
