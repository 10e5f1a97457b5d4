"""GPU diagnostic: where do tensor-core-engine gradients differ from fp32-engine gradients (same library, same inputs)?
Runs forward/backward with each engine combination and prints per-parameter-tensor max-relative and L2-relative errors."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, os.path.join(R, "cat-generator_b200")]
import numpy as np
from catgen import lib, models
from oracle import pyoracle as po
lib.init(0); L = lib.load()

def regions_G(kind, C):
    spec = [("lin.W", 8192 * 100), ("lin.b", 8192), ("prelu0", 1)]
    st = [(512, 512, 3, 1), (512, 256, 3, 1), (256, 128, 5, 1), (128, C, 3, 0)] if kind == lib.G32UPC else [(128, 256, 5, 1), (256, 128, 5, 1), (128, C, 3, 0)]
    for i, (ci, co, k, bn) in enumerate(st):
        spec += [("conv%d.W" % (i + 1), co * ci * k * k), ("conv%d.b" % (i + 1), co)]
        if bn: spec += [("bn%d.g" % (i + 1), co), ("bn%d.b" % (i + 1), co), ("prelu%d" % (i + 1), 1)]
    return spec
def regions_D(C):
    spec = []
    def stn(tag, ch, S, nth):
        f = 16 * (S // 4) ** 2
        spec.extend([(tag + ".c1W", 16*ch*9), (tag + ".c1b", 16), (tag + ".c2W", 2304), (tag + ".c2b", 16), (tag + ".l1W", 64*f), (tag + ".l1b", 64), (tag + ".l2W", nth*64), (tag + ".l2b", nth)])
    stn("stn0", C, 32, 1); spec.extend([("t1.W", 64*C*9), ("t1.b", 64), ("t1.p", 1), ("t2.W", 36864), ("t2.b", 64), ("t2.p", 1)])
    for b in range(3):
        stn("stn%d" % (b+1), 64, 16, 4); spec.extend([("br%d.c1W" % b, 36864), ("br%d.c1b" % b, 64), ("br%d.p1" % b, 1), ("br%d.c2W" % b, 36864), ("br%d.c2b" % b, 64), ("br%d.p2" % b, 1)])
    spec.extend([("br3.c1W", 128*64*25), ("br3.c1b", 128), ("br3.p1", 1), ("br3.c2W", 128*128*49), ("br3.c2b", 128), ("br3.p2", 1), ("h1.W", 256*20480), ("h1.b", 256), ("h.p", 1), ("h2.W", 256), ("h2.b", 1)])
    return spec
def report(title, g, ref, spec, only_big=True):
    gm = np.abs(ref).max(); o = 0
    print("  %s: overall max-rel %.2e  L2-rel %.2e" % (title, np.abs(g - ref).max() / gm, np.linalg.norm(g - ref) / np.linalg.norm(ref)))
    rows = []
    for n, k in spec:
        a, b = g[o:o+k], ref[o:o+k]; o += k
        rows.append((np.abs(a - b).max() / gm, n, np.abs(a - b).max() / (np.abs(b).max() + 1e-30), np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30), np.abs(b).max()))
    assert o == g.size
    for e, n, own, l2, mx in sorted(rows, reverse=True)[:6]:
        print("      %-10s err/globalmax %.2e   err/ownmax %.2e   L2-rel %.2e   |ref|max %.2e" % (n, e, own, l2, mx))

rng = np.random.default_rng(1)
# ---------------- G
for kind, okind, C in ((lib.G32UPC, po.G32UPC, 3),):
    B = 8
    og = po.Model(okind, C, 100, seed=1)
    g = models.create_G((C, 32, 32), 100, kind=kind); g.set_params(og.params); g.set_bn_running(og.bn_running)
    z = rng.uniform(-1, 1, (B, 100)).astype(np.float32); gout = (rng.standard_normal((B, C, 32, 32)) * 0.01).astype(np.float32)
    res = {}
    for fe, be in ((0, 0), (1, 1), (0, 1), (1, 0)):
        g.set_bn_running(og.bn_running)
        lib.check(L.cg_set_conv_engine(fe)); px = g.forward(z)
        lib.check(L.cg_set_conv_engine(be)); g.zeroGradParameters(); gz = g.backward(z, gout)
        res[(fe, be)] = (px.copy(), gz.copy(), g.get_grads())
    print("G32up-c B=8 (engine 0 = fp32 CUDA cores, 1 = tcgen05); reference = fwd fp32 / bwd fp32")
    ref = res[(0, 0)]
    for key, name in (((1, 1), "fwd TC / bwd TC  "), ((0, 1), "fwd fp32 / bwd TC"), ((1, 0), "fwd TC / bwd fp32")):
        r = res[key]
        print(" %s  pixels %.2e   gz max-rel %.2e L2-rel %.2e" % (name, np.abs(r[0] - ref[0]).max(), np.abs(r[1] - ref[1]).max() / np.abs(ref[1]).max(), np.linalg.norm(r[1] - ref[1]) / np.linalg.norm(ref[1])))
        report("param grads", r[2], ref[2], regions_G(kind, C))
# ---------------- D
B, C = 6, 3
od = po.Model(po.D32_ST3, C, 100, seed=3); p = od.params; p += rng.standard_normal(p.size).astype(np.float32) * 0.01
d = models.create_D((C, 32, 32), True); d.set_params(od.params); d.evaluate()
x = rng.uniform(0, 1, (B, C, 32, 32)).astype(np.float32); gout = rng.standard_normal(B).astype(np.float32)
res = {}
for fe, be in ((0, 0), (1, 1), (0, 1), (1, 0)):
    lib.check(L.cg_set_conv_engine(fe)); out, pre = d.forward(x, with_pre=True)
    lib.check(L.cg_set_conv_engine(be)); d.zeroGradParameters(); gx = d.backward(x, gout)
    res[(fe, be)] = (pre.copy(), gx.copy(), d.get_grads())
print("D32_st3 B=6 eval mode; reference = fwd fp32 / bwd fp32")
ref = res[(0, 0)]
for key, name in (((1, 1), "fwd TC / bwd TC  "), ((0, 1), "fwd fp32 / bwd TC"), ((1, 0), "fwd TC / bwd fp32")):
    r = res[key]
    e = np.abs(r[1] - ref[1]); gm = np.abs(ref[1]).max()
    print(" %s  pre-sigmoid %.2e   gx max-rel %.2e L2-rel %.2e  frac(|err| > 1e-2 max) %.5f" % (name, np.abs(r[0] - ref[0]).max(), e.max() / gm, np.linalg.norm(r[1] - ref[1]) / np.linalg.norm(ref[1]), np.mean(e > 1e-2 * gm)))
    report("param grads", r[2], ref[2], regions_D(C))
