"""Per-(layer, pass) table of every convolution shape the three networks instantiate, at one batch size, through the op-level C-ABI
with CUDA events around every launch (cg_profile_enable): us per launch and TFLOP/s of the tensor-core kernels.
    python tools/conv_layers.py [B] > profiles/r02_conv_layers.txt
The ncu columns (tensor-pipe %, xbar bytes, dram bytes) are added from a separate `ncu` run of tools/run_layer.py."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cat-generator_b200")]
from catgen import lib

LAYERS = [  # name, Ci, H, Co, k   (square images; reference file:line in models.lua)
    ("G32upc.conv1  models.lua:206", 512, 8, 512, 3), ("G32upc.conv2  :212", 512, 16, 256, 3), ("G32upc.conv3  :218", 256, 32, 128, 5),
    ("G32upc.conv4  :222", 128, 32, 3, 3), ("G32up.conv1   :145", 128, 16, 256, 5), ("G32up.conv2   :150", 256, 32, 128, 5),
    ("D.trunk1      :646", 3, 32, 64, 3), ("D.trunk2      :648", 64, 32, 64, 3), ("D.br1-3.conv1 :655", 64, 16, 64, 3),
    ("D.br1-3.conv2 :659", 64, 8, 64, 3), ("D.br4.conv1   :680", 64, 16, 128, 5), ("D.br4.conv2   :685", 128, 8, 128, 7),
    ("STN0.conv1    :844", 3, 16, 16, 3), ("STN0.conv2    :846", 16, 16, 16, 3), ("STN1-3.conv1  :844", 64, 8, 16, 3), ("STN1-3.conv2  :846", 16, 8, 16, 3),
]

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    only = sys.argv[2] if len(sys.argv) > 2 else None
    L = lib.load(); lib.init(0)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    rng = np.random.default_rng(0)
    print("batch %d; us per launch (mean of 5, CUDA events on the launching stream), TFLOP/s on algorithmic FLOPs 2*N*H*W*Co*Ci*k*k" % B)
    print("%-30s %-6s %-12s %9s %9s" % ("layer", "pass", "kernel", "us", "TFLOP/s"))
    for name, Ci, H, Co, k in LAYERS:
        if only and only not in name: continue
        N, W = B, H
        x = rng.uniform(-1, 1, (N, Ci, H, W)).astype(np.float32)
        Wt = (rng.uniform(-1, 1, (Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
        b = np.zeros(Co, np.float32); gy = (rng.standard_normal((N, Co, H, W)) * 1e-3).astype(np.float32)
        y = np.empty((N, Co, H, W), np.float32); gx = np.empty_like(x); gW = np.zeros_like(Wt); gb = np.zeros(Co, np.float32)
        passes = {
            "fprop": lambda: lib.check(L.cg_conv2d_fprop(P(x), P(Wt), P(b), P(y), N, Ci, H, W, Co, k)),
            "dgrad": lambda: lib.check(L.cg_conv2d_dgrad(P(gy), P(Wt), P(gx), N, Ci, H, W, Co, k)),
            "wgrad": lambda: lib.check(L.cg_conv2d_wgrad(P(x), P(gy), P(gW), P(gb), N, Ci, H, W, Co, k)),
        }
        fl = 2.0 * N * H * W * Co * Ci * k * k
        for pname, fn in passes.items():
            fn(); fn()
            lib.check(L.cg_profile_enable(1))
            for _ in range(5): fn()
            buf = C.create_string_buffer(1 << 16)
            lib.check(L.cg_profile_report(buf, len(buf))); lib.check(L.cg_profile_enable(0))
            for r in json.loads(buf.value.decode()):
                if r["kernel"].startswith(("k_conv_tc", "k_conv_ps", "k_wgrad")):
                    us = 1e3 * r["ms"] / r["launches"]
                    print("%-30s %-6s %-12s %9.1f %9.1f" % (name, pname, r["kernel"][:12], us, fl / us / 1e6), flush=True)

if __name__ == "__main__":
    main()
