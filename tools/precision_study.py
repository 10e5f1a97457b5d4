"""CPU estimate of generated-pixel error when conv operands are rounded to tf32/fp16/bf16 (fp32 accumulate),
against a float64 PyTorch restatement of G32up-c.  Models operand rounding only; init-scale weights.
Off the product path.  Result of the run that fixed the precision choice is in profiles/r01_precision_study.txt."""
import sys, time, numpy as np, torch, torch.nn.functional as F
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, os.path.join(R, "tests")]
from oracle import pyoracle as po
from oracle import torch_ref as tr
torch.set_num_threads(8)
def rn_tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
ROUND = {"fp32": lambda x: x, "tf32": rn_tf32, "fp16": lambda x: x.half().float(), "bf16": lambda x: x.bfloat16().float()}
def G(flat, z, rnd, round_last):
    C0, s0, stages = tr.g_spec("G32UPC", 3); c = tr.Cursor(flat)
    W, b, pw = c.take(C0*s0*s0, 100), c.take(C0*s0*s0), c.take(1)
    x = F.prelu(F.linear(z, W, b), pw).view(-1, C0, s0, s0)
    for i, (up, Ci, Co, k, bn) in enumerate(stages):
        if up: x = F.interpolate(x, scale_factor=2, mode="nearest")
        W, b = c.take(Co, Ci, k, k), c.take(Co)
        r = rnd if (bn or round_last) else (lambda t: t)
        x = F.conv2d(r(x), r(W), b, padding=(k-1)//2)
        if bn:
            g, bt, pw = c.take(Co), c.take(Co), c.take(1)
            x = F.prelu(F.batch_norm(x, None, None, g, bt, training=True, eps=1e-5), pw)
        else: x = torch.sigmoid(x)
    return x
B = 64
g = po.Model(po.G32UPC, 3, 100, seed=1)
rng = np.random.default_rng(1)
z = torch.tensor(rng.uniform(-1, 1, (B, 100)).astype(np.float32))
flat = torch.tensor(g.params.copy())
t0 = time.time()
with torch.no_grad():
    truth = tr.G_forward(flat.double(), z.double(), "G32UPC", 3).numpy()
    print("truth f64 done %.0fs; pixel range [%.3f, %.3f]" % (time.time()-t0, truth.min(), truth.max()))
    for name in ("fp32", "tf32", "fp16", "bf16"):
        for rl in ((False,) if name == "fp32" else (False, True)):
            o = G(flat, z, ROUND[name], rl).numpy().astype(np.float64)
            e = np.abs(o - truth)
            print("%-5s conv4 %-8s max-abs %.2e  p99.9 %.2e  mean %.2e  -> %s 1e-3" % (
                name, "rounded" if rl else "fp32", e.max(), np.quantile(e, 0.999), e.mean(),
                "WITHIN" if e.max() < 1e-3 else "EXCEEDS"))
