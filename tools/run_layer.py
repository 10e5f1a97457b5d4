"""Run ONE convolution (fprop | dgrad | wgrad) through the C-ABI a few times, so `ncu -k regex:<kernel>` can capture it
in isolation.  usage: run_layer.py <fprop|dgrad|wgrad> N Ci H W Co k [reps]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, os.path.join(R, "cat-generator_b200")]
import numpy as np
from catgen import lib
op = sys.argv[1]; N, Ci, H, W, Co, k = map(int, sys.argv[2:8]); reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
lib.init(0); L = lib.load(); P = lib.P
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, (N, Ci, H, W)).astype(np.float32)
Wt = (rng.uniform(-1, 1, (Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
b = np.zeros(Co, np.float32)
gy = (rng.standard_normal((N, Co, H, W)) * 1e-4).astype(np.float32)
y = np.empty((N, Co, H, W), np.float32); gx = np.empty_like(x); gW = np.zeros_like(Wt); gb = np.zeros(Co, np.float32)
for _ in range(reps):
    if op == "fprop": lib.check(L.cg_conv2d_fprop(P(x), P(Wt), P(b), P(y), N, Ci, H, W, Co, k))
    elif op == "dgrad": lib.check(L.cg_conv2d_dgrad(P(gy), P(Wt), P(gx), N, Ci, H, W, Co, k))
    else: lib.check(L.cg_conv2d_wgrad(P(x), P(gy), P(gW), P(gb), N, Ci, H, W, Co, k))
print("done", op, np.isfinite(y).all(), np.isfinite(gx).all(), np.isfinite(gW).all())
