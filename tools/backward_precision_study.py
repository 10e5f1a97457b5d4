"""How well-conditioned is G's BACKWARD pass in a real GAN step, and what does operand rounding in the backward
convolutions do to G's parameter gradient?  Off the product path; CPU only.

Truth = float64 PyTorch restatement of G (oracle/torch_ref.py) back-propagating the ACTUAL image gradient that D
produces for the generator update (oracle fevalG_on_D path), not random noise.  Compared against it:
  fp32            : the same graph in float32 (an honest fp32 implementation)
  tf32/fp16/bf16/fp16-scaled : float64 arithmetic, but the operands of every conv dgrad/wgrad (gy, W, x) rounded to that format
                    (models tensor-core operand rounding with fp32+ accumulation); forward kept exact
Also reports the conditioning of the batch-norm backward: |gx| / |g_in| per BN layer.
"""
import os, sys
import numpy as np, torch, torch.nn.functional as F
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, os.path.join(R, "tests")]
from oracle import pyoracle as po
from oracle import torch_ref as tr
torch.set_num_threads(8)

def rn_tf32(x):
    x32 = x.float().contiguous(); i = x32.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).to(x.dtype)
def fp16_scaled(x):
    """fp16 with a per-tensor power-of-two scale putting max|x| near 2^14 (what a max-reduction + pack kernel does)."""
    m = float(x.abs().max())
    if m == 0: return x
    import math
    sc = 2.0 ** math.floor(math.log2(16384.0 / m))
    return (x * sc).half().to(x.dtype) / sc
ROUND = {"tf32": rn_tf32, "fp16": lambda x: x.half().to(x.dtype), "bf16": lambda x: x.bfloat16().to(x.dtype), "fp16-scaled": fp16_scaled}

class ConvR(torch.autograd.Function):
    """conv2d whose BACKWARD rounds its operands (forward exact)."""
    @staticmethod
    def forward(ctx, x, W, b, pad, rnd):
        ctx.save_for_backward(x, W); ctx.pad, ctx.rnd = pad, rnd
        return F.conv2d(x, W, b, padding=pad)
    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors; r = ctx.rnd
        gyr, Wr, xr = r(gy), r(W), r(x)
        gx = torch.nn.grad.conv2d_input(x.shape, Wr, gyr, padding=ctx.pad)
        gW = torch.nn.grad.conv2d_weight(xr, W.shape, gyr, padding=ctx.pad)
        return gx, gW, gy.sum((0, 2, 3)), None, None

def G_forward(flat, z, kind, C, rnd, taps):
    C0, s0, stages = tr.g_spec(kind, C); c = tr.Cursor(flat)
    W, b, pw = c.take(C0*s0*s0, 100), c.take(C0*s0*s0), c.take(1)
    x = F.prelu(F.linear(z, W, b), pw).view(-1, C0, s0, s0)
    for up, Ci, Co, k, bn in stages:
        if up: x = F.interpolate(x, scale_factor=2, mode="nearest")
        W, b = c.take(Co, Ci, k, k), c.take(Co)
        x = ConvR.apply(x, W, b, (k-1)//2, rnd) if rnd else F.conv2d(x, W, b, padding=(k-1)//2)
        if bn:
            g, bt, pw = c.take(Co), c.take(Co), c.take(1)
            x.retain_grad() if x.requires_grad else None; taps.append(x)
            y = F.batch_norm(x, None, None, g, bt, training=True, eps=1e-5); y.retain_grad(); taps.append(y)
            x = F.prelu(y, pw)
        else: x = torch.sigmoid(x)
    return x

def rel(a, b): return float((a - b).abs().max() / b.abs().max())

def study(kind, okind, C, B, seed=5):
    rng = np.random.default_rng(seed)
    og, od = po.Model(okind, C, 100, seed=1), po.Model(po.D32_ST3, C, 100, seed=2)
    z = rng.uniform(-1, 1, (B, 100)).astype(np.float32)
    px = og.G_forward(z, True); masks = po.make_D_masks(B, rng)
    sig, _ = od.D_forward(px, masks)
    df = np.empty(B, np.float32); po.lib().og_bce_bwd(po.P(sig), po.P(np.ones(B, np.float32)), po.P(df), B)
    gimg = od.D_backward(df)                                   # the image gradient the generator update receives
    print("== %s C=%d B=%d | D out %.3f..%.3f | |gimg| max %.2e, common-mode fraction %.3f" % (
        kind, C, B, sig.min(), sig.max(), np.abs(gimg).max(), np.abs(gimg.mean(0)).max() / np.abs(gimg).max()))
    res = {}
    for name in ("f64", "fp32", "tf32", "fp16-scaled", "fp16", "bf16"):
        dt = torch.float32 if name == "fp32" else torch.float64
        flat = torch.tensor(og.params.copy()).to(dt).requires_grad_(); zt = torch.tensor(z).to(dt)
        taps = []
        out = G_forward(flat, zt, kind, C, ROUND.get(name), taps)
        out.backward(torch.tensor(gimg).to(dt))
        res[name] = flat.grad.double()
        if name == "f64":
            for i in range(0, len(taps), 2):
                print("   BN%d backward: |g_in| %.2e -> |g_out| %.2e  (ratio %.1e: cancellation amplifies relative error by ~1/ratio)" % (
                    i // 2 + 1, float(taps[i+1].grad.abs().max()), float(taps[i].grad.abs().max()), float(taps[i].grad.abs().max() / taps[i+1].grad.abs().max())))
    L = po.lib()
    for t in (1, 2):
        L.og_set_threads(t); og.zero_grads(); og.G_forward(z, True); og.G_backward(gimg)
        res["oracle(%d thr)" % t] = torch.tensor(og.grads.copy()).double()
    for name, g in res.items():
        if name != "f64": print("   %-14s param-grad error vs f64, rel-to-max: %.2e" % (name, rel(g, res["f64"])))

if __name__ == "__main__":
    study("G32UP", po.G32UP, 1, 8)        # the configuration of the failing train-step test
    study("G32UPC", po.G32UPC, 3, 8)
    study("G32UPC", po.G32UPC, 3, 64)     # c2-like batch (half of 128 to keep the CPU run short)
