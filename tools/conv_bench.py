"""Per-kernel time of one conv layer's three passes through the op-level C-ABI (profiling events around every launch).
Default shape: G32up-c conv3 at the BASELINE batch (128 x 256 x 32 x 32 -> 128 channels, 5x5).  Used with the engine's
experiment knobs (CATGEN_TC_ROT, CATGEN_TC_RING) to see what the tensor-core kernels are sensitive to.
    python tools/conv_bench.py [N Ci H W Co k]"""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cat-generator_b200")]
from catgen import lib

def main():
    shape = [int(a) for a in sys.argv[1:7]] if len(sys.argv) >= 7 else [128, 256, 32, 32, 128, 5]
    N, Ci, H, W, Co, k = shape
    L = lib.load(); lib.init(0)
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (N, Ci, H, W)).astype(np.float32)
    Wt = (rng.uniform(-1, 1, (Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
    b = np.zeros(Co, np.float32); gy = rng.standard_normal((N, Co, H, W)).astype(np.float32)
    y = np.empty((N, Co, H, W), np.float32); gx = np.empty_like(x); gW = np.zeros_like(Wt); gb = np.zeros(Co, np.float32)
    passes = {
        "fprop": lambda: lib.check(L.cg_conv2d_fprop(P(x), P(Wt), P(b), P(y), N, Ci, H, W, Co, k)),
        "dgrad": lambda: lib.check(L.cg_conv2d_dgrad(P(gy), P(Wt), P(gx), N, Ci, H, W, Co, k)),
        "wgrad": lambda: lib.check(L.cg_conv2d_wgrad(P(x), P(gy), P(gW), P(gb), N, Ci, H, W, Co, k)),
    }
    fl = 2.0 * N * H * W * Co * Ci * k * k
    tag = " ".join("%s=%s" % (e, os.environ[e]) for e in ("CATGEN_TC_ROT", "CATGEN_TC_RING") if e in os.environ) or "default"
    for name, fn in passes.items():
        fn(); fn()
        lib.check(L.cg_profile_enable(1))
        for _ in range(5): fn()
        buf = C.create_string_buffer(1 << 16)
        lib.check(L.cg_profile_report(buf, len(buf))); lib.check(L.cg_profile_enable(0))
        for r in json.loads(buf.value.decode()):
            if r["kernel"].startswith(("k_conv_tc", "k_conv_ps", "k_wgrad_tc")):
                us = 1e3 * r["ms"] / r["launches"]
                print("%-34s %-6s %-14s %2d launches  %8.1f us/launch  %7.1f TFLOP/s" % (tag, name, r["kernel"], r["launches"], us, fl / us / 1e6), flush=True)

if __name__ == "__main__":
    main()
