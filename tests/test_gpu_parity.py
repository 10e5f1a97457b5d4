"""GPU parity: libcatgen (through the C-ABI) against the CPU oracle on identical seeded inputs.

Tolerances (floating point; stated per check):
  * op level, fp32 engine: |a-b| <= 2e-4 * max|b| (different summation order only).
  * generated pixels: <= 1e-3 max-abs (BASELINE.json north_star); D sigmoid outputs <= 1e-3.
  * gradients, fp32 engine: 1e-2 of max|oracle| (SURVEY.md section 8c proposal; the oracle itself is ~1e-3 from a
    float64 restatement on G32up-c, profiles/r01_parity_noise_floor.txt).
  * gradients, tensor-core engine -- three separate statements (profiles/r01_grad_diag.txt):
      (a) the backward KERNELS are exact: with an fp32 forward and a tensor-core backward, gradients are within
          2e-3 of the fp32 engine (measured 2e-4 on G, 6e-6 on D);
      (b) the forward is within the north_star bound (pixels 1.8e-4, D pre-sigmoid 5e-6);
      (c) end to end the gradient differs by ~2e-2 on G and ~3e-2 (L2) on D's input gradient, ALL of it caused by the
          fp16 forward moving pre-activations across PReLU / max-pool decision points: a forward perturbation eps
          flips ~eps of the decisions and leaves ~sqrt(eps) diffuse noise in sums over millions of elements (the
          same law gives the ~1e-3 oracle-vs-oracle figure at fp32 eps).  The 1e-2 proposal is therefore NOT met
          end to end with fp16 operands; the bound asserted is 5e-2 of max (G, D parameters) and 1e-1 in L2 for D's
          input gradient, whose max-abs is dominated by individual max-pool reroutes.
The oracle is the checker only; nothing under test calls it.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po          # noqa: E402
from catgen import lib, models, adversarial  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _init():
    lib.init(0)
    yield


@pytest.fixture(params=[0, 1], ids=["fp32-engine", "tc-engine"])
def engine(request):
    L = lib.load()
    lib.check(L.cg_set_conv_engine(request.param))
    yield request.param
    lib.check(L.cg_set_conv_engine(1))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


P = lib.P
OP_TOL = 2e-4


def l2rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300)


def gtol(engine):
    """end-to-end gradient bound relative to max|oracle| (module docstring, statement (c))"""
    return 1e-2 if engine == 0 else 5e-2


def test_gpu_launches_are_counted():
    L = lib.load()
    L.cg_reset_launch_count()
    x = np.ones(1000, np.float32); y = np.empty_like(x)
    lib.check(L.cg_sigmoid_fwd(P(x), P(y), 1000))
    assert L.cg_launch_count() >= 1
    assert np.allclose(y, 1 / (1 + np.exp(-1)), atol=1e-6)


# ------------------------------------------------------------------ convolution
CONV_SHAPES = [  # N, Ci, H, W, Co, k
    (2, 3, 32, 32, 64, 3),     # D trunk conv1 (models.lua:646)
    (2, 64, 32, 32, 64, 3),    # D trunk conv2 (:648)
    (2, 64, 16, 16, 128, 5),   # D branch 4 (:680)
    (2, 128, 8, 8, 128, 7),    # D branch 4 (:685)
    (2, 16, 8, 8, 16, 3),      # STN loc-net (:846)
    (2, 128, 32, 32, 3, 3),    # G conv4 (:222)
    (2, 512, 8, 8, 512, 3),    # G conv1 (:206)
    (1, 256, 32, 32, 128, 5),  # G conv3 (:218)
    (3, 5, 7, 9, 6, 3),        # ragged: nothing a multiple of anything
    (1, 1, 32, 32, 64, 3),     # grayscale (--colorSpace=y)
    (2, 4, 1, 1, 4, 1),        # degenerate 1x1
]


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_fprop_dgrad_wgrad(shape, engine):
    N, Ci, H, W, Co, k = shape
    rng = np.random.default_rng(hash(shape) % 2**31)
    L, O = lib.load(), po.lib()
    x = rng.uniform(-1, 1, (N, Ci, H, W)).astype(np.float32)
    Wt = (rng.uniform(-1, 1, (Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, Co).astype(np.float32)
    gy = rng.standard_normal((N, Co, H, W)).astype(np.float32)
    tol = OP_TOL if engine == 0 else 3e-3   # tensor-core path: fp16/tf32 operands, fp32 accumulate
    y, y0 = np.empty((N, Co, H, W), np.float32), np.empty((N, Co, H, W), np.float32)
    lib.check(L.cg_conv2d_fprop(P(x), P(Wt), P(b), P(y), N, Ci, H, W, Co, k))
    O.og_conv2d_fwd(po.P(x), po.P(Wt), po.P(b), po.P(y0), N, Ci, H, W, Co, k)
    assert rel(y, y0) < tol
    gx, gx0 = np.empty_like(x), np.empty_like(x)
    lib.check(L.cg_conv2d_dgrad(P(gy), P(Wt), P(gx), N, Ci, H, W, Co, k))
    O.og_conv2d_bwd_data(po.P(gy), po.P(Wt), po.P(gx0), N, Ci, H, W, Co, k)
    assert rel(gx, gx0) < tol
    # accGradParameters ACCUMULATES: start from a non-zero gradient
    gW = rng.standard_normal(Wt.shape).astype(np.float32); gb = rng.standard_normal(Co).astype(np.float32)
    gW0, gb0 = gW.copy(), gb.copy()
    lib.check(L.cg_conv2d_wgrad(P(x), P(gy), P(gW), P(gb), N, Ci, H, W, Co, k))
    O.og_conv2d_bwd_filter(po.P(x), po.P(gy), po.P(gW0), po.P(gb0), N, Ci, H, W, Co, k)
    assert rel(gW, gW0) < tol and rel(gb, gb0) < tol


def test_conv_upsample_module():
    """layers/SpatialConvolutionUpsample.lua: conv to nOut*f^2 planes + contiguous view; even k is an error."""
    L, O = lib.load(), po.lib()
    rng = np.random.default_rng(5)
    N, Ci, H, W, nOut, k, f = 2, 8, 6, 6, 4, 3, 2
    x = rng.uniform(-1, 1, (N, Ci, H, W)).astype(np.float32)
    Wt = rng.uniform(-0.2, 0.2, (nOut * f * f, Ci, k, k)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, nOut * f * f).astype(np.float32)
    y, y0 = np.empty((N, nOut, H * f, W * f), np.float32), np.empty((N, nOut * f * f, H, W), np.float32)
    lib.check(L.cg_conv_upsample_fwd(P(x), P(Wt), P(b), P(y), N, Ci, H, W, nOut, k, f))
    O.og_conv2d_fwd(po.P(x), po.P(Wt), po.P(b), po.P(y0), N, Ci, H, W, nOut * f * f, k)
    assert rel(y, y0.reshape(y.shape)) < OP_TOL      # the Lua :view(...) reinterprets the same memory
    # index map stated in SURVEY.md section 8 row A9: plane p fills output rows [p*h/2, (p+1)*h/2) for f=2
    assert np.array_equal(y[0, 0, :H // 2].reshape(-1), y.reshape(N, nOut * f * f, H, W)[0, 0].reshape(-1)[:H // 2 * W * f])
    gy = rng.standard_normal(y.shape).astype(np.float32)
    gx, gW, gb = np.empty_like(x), np.zeros_like(Wt), np.zeros_like(b)
    lib.check(L.cg_conv_upsample_bwd(P(x), P(gy), P(Wt), P(gx), P(gW), P(gb), N, Ci, H, W, nOut, k, f))
    gx0, gW0, gb0 = np.empty_like(x), np.zeros_like(Wt), np.zeros_like(b)
    O.og_conv2d_bwd_data(po.P(gy), po.P(Wt), po.P(gx0), N, Ci, H, W, nOut * f * f, k)
    O.og_conv2d_bwd_filter(po.P(x), po.P(gy), po.P(gW0), po.P(gb0), N, Ci, H, W, nOut * f * f, k)
    assert rel(gx, gx0) < OP_TOL and rel(gW, gW0) < OP_TOL and rel(gb, gb0) < OP_TOL
    assert L.cg_conv_upsample_fwd(P(x), P(Wt), P(b), P(y), N, Ci, H, W, nOut, 4, f) != 0   # assert(kW % 2 == 1)
    assert b"odd" in L.cg_last_error()


@pytest.mark.parametrize("N,inn,out", [(8, 100, 8192), (4, 20480, 256), (5, 256, 1), (3, 64, 4), (1, 7, 3),
                                       (128, 20480, 256), (64, 100, 8192), (256, 256, 1), (128, 1024, 64), (128, 64, 4)],
                         ids=lambda v: str(v))
def test_linear(N, inn, out, engine):
    """nn.Linear (models.lua:199,697,700,852,854).  On the tensor-core engine batches that are a multiple of 8 (<= 128) or of 128 run
    forward and input gradient as a 1x1 convolution over the batch rows on tcgen05, split along K when the tile grid is small
    (csrc/conv_tc.cu linear_tc_run); fp16 operands, fp32 accumulate: 3e-3 of max.  The weight gradient stays on the fp32 kernel."""
    L, O = lib.load(), po.lib()
    rng = np.random.default_rng(N * 1000 + out)
    x = rng.uniform(-1, 1, (N, inn)).astype(np.float32)
    W = (rng.uniform(-1, 1, (out, inn)) / np.sqrt(inn)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, out).astype(np.float32)
    gy = rng.standard_normal((N, out)).astype(np.float32)
    tol = OP_TOL if (engine == 0 or N % 8) else 3e-3
    y, y0 = np.empty((N, out), np.float32), np.empty((N, out), np.float32)
    lib.check(L.cg_linear_fwd(P(x), P(W), P(b), P(y), N, inn, out)); O.og_linear_fwd(po.P(x), po.P(W), po.P(b), po.P(y0), N, inn, out)
    assert rel(y, y0) < tol
    gx, gW, gb = np.empty_like(x), np.ones_like(W), np.ones_like(b)
    gx0, gW0, gb0 = np.empty_like(x), np.ones_like(W), np.ones_like(b)
    lib.check(L.cg_linear_bwd(P(x), P(gy), P(W), P(gx), P(gW), P(gb), N, inn, out))
    O.og_linear_bwd(po.P(x), po.P(gy), po.P(W), po.P(gx0), po.P(gW0), po.P(gb0), N, inn, out)
    assert rel(gx, gx0) < tol and rel(gW, gW0) < OP_TOL and rel(gb, gb0) < OP_TOL


@pytest.mark.parametrize("N,Cc,HW", [(8, 512, 64), (4, 128, 1024), (3, 5, 7), (2, 1, 4)])
def test_batchnorm(N, Cc, HW):
    L, O = lib.load(), po.lib()
    rng = np.random.default_rng(Cc)
    x = (rng.standard_normal((N, Cc, HW)) * 2 + 0.5).astype(np.float32)
    g = rng.uniform(0, 1, Cc).astype(np.float32); bt = rng.uniform(-0.5, 0.5, Cc).astype(np.float32)
    gy = rng.standard_normal(x.shape).astype(np.float32)
    out = {}
    for name in ("gpu", "cpu"):
        y = np.empty_like(x); m = np.empty(Cc, np.float32); iv = np.empty(Cc, np.float32)
        rm = np.zeros(Cc, np.float32); rv = np.ones(Cc, np.float32)
        gx = np.empty_like(x); gg = np.zeros(Cc, np.float32); gb = np.zeros(Cc, np.float32)
        if name == "gpu":
            lib.check(L.cg_bn2d_fwd(P(x), P(g), P(bt), P(y), P(m), P(iv), P(rm), P(rv), N, Cc, HW))
            lib.check(L.cg_bn2d_bwd(P(x), P(gy), P(g), P(m), P(iv), P(gx), P(gg), P(gb), N, Cc, HW))
        else:
            O.og_bn_fwd_train(po.P(x), po.P(g), po.P(bt), po.P(y), po.P(m), po.P(iv), po.P(rm), po.P(rv), N, Cc, HW, 1e-5, 0.1)
            O.og_bn_bwd_train(po.P(x), po.P(gy), po.P(g), po.P(m), po.P(iv), po.P(gx), po.P(gg), po.P(gb), N, Cc, HW)
        out[name] = (y, m, iv, rm, rv, gx, gg, gb)
    for a, b in zip(out["gpu"], out["cpu"]):
        assert rel(a, b) < OP_TOL


def test_pointwise_pool_upsample():
    L, O = lib.load(), po.lib()
    rng = np.random.default_rng(11)
    n = 100003
    x = rng.standard_normal(n).astype(np.float32); x[:5] = 0.0          # exact zeros: LeakyReLU grad at 0 is gy
    gy = rng.standard_normal(n).astype(np.float32)
    for fwd, bwd, ofwd, obwd, arg in (("cg_leakyrelu_fwd", "cg_leakyrelu_bwd", "og_leakyrelu_fwd", "og_leakyrelu_bwd", 0.333),):
        y, y0, gx, gx0 = (np.empty(n, np.float32) for _ in range(4))
        lib.check(getattr(L, fwd)(P(x), arg, P(y), n)); getattr(O, ofwd)(po.P(x), arg, po.P(y0), n)
        lib.check(getattr(L, bwd)(P(x), P(gy), arg, P(gx), n)); getattr(O, obwd)(po.P(x), po.P(gy), arg, po.P(gx0), n)
        assert np.array_equal(y, y0) and np.array_equal(gx, gx0)
        assert np.array_equal(gx[:5], gy[:5])
    y, y0, gx, gx0 = (np.empty(n, np.float32) for _ in range(4)); gw = np.array([0.5], np.float32); gw0 = gw.copy()
    lib.check(L.cg_prelu_fwd(P(x), 0.25, P(y), n)); O.og_prelu_fwd(po.P(x), 0.25, po.P(y0), n)
    lib.check(L.cg_prelu_bwd(P(x), P(gy), 0.25, P(gx), P(gw), n)); O.og_prelu_bwd(po.P(x), po.P(gy), 0.25, po.P(gx0), po.P(gw0), n)
    assert np.array_equal(y, y0) and np.array_equal(gx, gx0) and rel(gw, gw0) < 1e-5
    lib.check(L.cg_sigmoid_fwd(P(x), P(y), n)); O.og_sigmoid_fwd(po.P(x), po.P(y0), n)
    assert np.abs(y - y0).max() < 1e-6
    lib.check(L.cg_sigmoid_bwd(P(y0), P(gy), P(gx), n)); O.og_sigmoid_bwd(po.P(y0), po.P(gy), po.P(gx0), n)
    assert rel(gx, gx0) < 1e-6
    NC, H, W = 7, 6, 10
    a = rng.standard_normal((NC, H, W)).astype(np.float32); a[0, 0, 0] = a[0, 0, 1] = 3.0   # tie: first max wins
    up, up0 = np.empty((NC, 2 * H, 2 * W), np.float32), np.empty((NC, 2 * H, 2 * W), np.float32)
    lib.check(L.cg_upsample2x_fwd(P(a), P(up), NC, H, W)); O.og_upsample2x_fwd(po.P(a), po.P(up0), NC, H, W)
    assert np.array_equal(up, up0)
    g2 = rng.standard_normal(up.shape).astype(np.float32); d, d0 = np.empty_like(a), np.empty_like(a)
    lib.check(L.cg_upsample2x_bwd(P(g2), P(d), NC, H, W)); O.og_upsample2x_bwd(po.P(g2), po.P(d0), NC, H, W)
    assert rel(d, d0) < 1e-6
    pl, pl0 = np.empty((NC, H // 2, W // 2), np.float32), np.empty((NC, H // 2, W // 2), np.float32)
    lib.check(L.cg_avgpool2_fwd(P(a), P(pl), NC, H, W)); O.og_avgpool2_fwd(po.P(a), po.P(pl0), NC, H, W)
    assert rel(pl, pl0) < 1e-6
    gp = rng.standard_normal(pl.shape).astype(np.float32)
    lib.check(L.cg_avgpool2_bwd(P(gp), P(d), NC, H, W)); O.og_avgpool2_bwd(po.P(gp), po.P(d0), NC, H, W)
    assert np.array_equal(d, d0)
    idx, idx0 = np.empty(pl.shape, np.int32), np.empty(pl.shape, np.int32)
    lib.check(L.cg_maxpool2_fwd(P(a), P(pl), idx.ctypes.data_as(C.POINTER(C.c_int32)), NC, H, W)); O.og_maxpool2_fwd(po.P(a), po.P(pl0), po.IP(idx0), NC, H, W)
    assert np.array_equal(pl, pl0) and np.array_equal(idx, idx0) and idx[0, 0, 0] == 0
    lib.check(L.cg_maxpool2_bwd(P(gp), idx.ctypes.data_as(C.POINTER(C.c_int32)), P(d), NC, H, W)); O.og_maxpool2_bwd(po.P(gp), po.IP(idx0), po.P(d0), NC, H, W)
    assert np.array_equal(d, d0)


@pytest.mark.parametrize("rot,scl,trn", [(1, 0, 0), (1, 1, 1)])
def test_spatial_transformer_pieces(rot, scl, trn):
    L, O = lib.load(), po.lib()
    rng = np.random.default_rng(3 + scl)
    B, S, Cc = 5, 16, 7
    nth = rot + scl + 2 * trn
    th = (rng.standard_normal((B, nth)) * 0.3).astype(np.float32)
    if scl: th[:, 1] += 1.0
    A, A0 = np.empty((B, 6), np.float32), np.empty((B, 6), np.float32)
    lib.check(L.cg_affine_matrix_fwd(P(th), P(A), B, rot, scl, trn)); O.og_affine_matrix_fwd(po.P(th), po.P(A0), B, rot, scl, trn)
    assert np.abs(A - A0).max() < 1e-6
    gA = rng.standard_normal((B, 6)).astype(np.float32); gt, gt0 = np.empty_like(th), np.empty_like(th)
    lib.check(L.cg_affine_matrix_bwd(P(th), P(gA), P(gt), B, rot, scl, trn)); O.og_affine_matrix_bwd(po.P(th), po.P(gA), po.P(gt0), B, rot, scl, trn)
    assert rel(gt, gt0) < 1e-5
    grid, grid0 = np.empty((B, S, S, 2), np.float32), np.empty((B, S, S, 2), np.float32)
    lib.check(L.cg_affine_grid_fwd(P(A0), P(grid), B, S, S)); O.og_affine_grid_fwd(po.P(A0), po.P(grid0), B, S, S)
    assert np.abs(grid - grid0).max() < 1e-6
    gg = rng.standard_normal(grid.shape).astype(np.float32); g6, g60 = np.empty((B, 6), np.float32), np.empty((B, 6), np.float32)
    lib.check(L.cg_affine_grid_bwd(P(gg), P(g6), B, S, S)); O.og_affine_grid_bwd(po.P(gg), po.P(g60), B, S, S)
    assert rel(g6, g60) < 1e-5
    img = rng.standard_normal((B, S, S, Cc)).astype(np.float32)
    out, out0 = np.empty_like(img), np.empty_like(img)
    lib.check(L.cg_bilinear_fwd(P(img), P(grid0), P(out), B, S, S, Cc)); O.og_bilinear_fwd(po.P(img), po.P(grid0), po.P(out0), B, S, S, Cc)
    assert np.abs(out - out0).max() < 1e-5       # includes out-of-range corners (translation pushes samples outside)
    go = rng.standard_normal(img.shape).astype(np.float32)
    gi, gi0, gr, gr0 = np.empty_like(img), np.empty_like(img), np.empty_like(grid), np.empty_like(grid)
    lib.check(L.cg_bilinear_bwd(P(img), P(grid0), P(go), P(gi), P(gr), B, S, S, Cc))
    O.og_bilinear_bwd(po.P(img), po.P(grid0), po.P(go), po.P(gi0), po.P(gr0), B, S, S, Cc)
    assert rel(gi, gi0) < 1e-5 and rel(gr, gr0) < 1e-4     # scatter-add uses fp32 atomics: order-dependent rounding only


def test_bce_and_edge_values():
    L, O = lib.load(), po.lib()
    p = np.array([0.1, 0.9, 0.5, 1e-7, 1 - 1e-7, 0.0, 1.0], np.float32)      # saturated D outputs included
    t = np.array([0, 1, 1, 0, 1, 0, 1], np.float32)
    loss = np.zeros(1, np.float32); g, g0 = np.empty(7, np.float32), np.empty(7, np.float32)
    lib.check(L.cg_bce(P(p), P(t), 7, P(loss), P(g)))
    l0 = O.og_bce_fwd(po.P(p), po.P(t), 7); O.og_bce_bwd(po.P(p), po.P(t), po.P(g0), 7)
    assert abs(loss[0] - l0) < 1e-6 and rel(g, g0) < 1e-5


# ------------------------------------------------------------------ models
G_CASES = [(lib.G32UPC, po.G32UPC, 3), (lib.G32UP, po.G32UP, 1), (lib.G32UP, po.G32UP, 3)]


@pytest.mark.parametrize("kind,okind,Cc", G_CASES, ids=["G32up-c-rgb", "G32up-gray", "G32up-rgb"])
def test_G_forward_backward(kind, okind, Cc, engine):
    rng = np.random.default_rng(1)
    B = 8
    og = po.Model(okind, Cc, 100, seed=1)
    g = models.create_G((Cc, 32, 32), 100, seed=7, kind=kind)
    assert g.nparams == og.n
    g.set_params(og.params)
    g.set_bn_running(og.bn_running)
    z = rng.uniform(-1, 1, (B, 100)).astype(np.float32)
    gout = (rng.standard_normal((B, Cc, 32, 32)) * 0.01).astype(np.float32)
    px0 = og.G_forward(z, train=True)
    px = g.forward(z)
    assert np.abs(px - px0).max() < 1e-3, "generated pixels (north_star tolerance)"
    if engine == 0:
        assert np.abs(px - px0).max() < 2e-5
    assert rel(g.get_bn_running(), og.bn_running) < (1e-4 if engine == 0 else 5e-3)
    og.zero_grads(); gz0 = og.G_backward(gout)
    g.zeroGradParameters(); gz = g.backward(z, gout)
    assert rel(gz, gz0) < gtol(engine) and rel(g.get_grads(), og.grads) < gtol(engine)
    # gradients accumulate across backward calls (accGradParameters)
    g.forward(z); g.backward(z, gout)
    assert rel(g.get_grads(), 2 * og.grads) < gtol(engine)


def test_tc_backward_kernels_exact_given_fp32_forward():
    """Statement (a): forward on the fp32 engine, backward on the tensor cores (dgrad tf32, wgrad scaled fp16, both
    MN-/K-major layouts verified by tools/tc_probe.cu).  Same saved activations => no decision point can flip, so
    the gradients must agree with the all-fp32 run to operand-rounding level and with the oracle to its own floor."""
    L = lib.load()
    rng = np.random.default_rng(1)
    try:
        # G32up-c
        B, Cc = 8, 3
        og = po.Model(po.G32UPC, Cc, 100, seed=1)
        g = models.create_G((Cc, 32, 32), 100); g.set_params(og.params)
        z = rng.uniform(-1, 1, (B, 100)).astype(np.float32); gout = (rng.standard_normal((B, Cc, 32, 32)) * 0.01).astype(np.float32)
        res = {}
        for be in (0, 1):
            g.set_bn_running(og.bn_running)
            lib.check(L.cg_set_conv_engine(0)); g.forward(z)
            lib.check(L.cg_set_conv_engine(be)); g.zeroGradParameters(); gz = g.backward(z, gout)
            res[be] = (gz.copy(), g.get_grads())
        assert rel(res[1][0], res[0][0]) < 2e-3 and rel(res[1][1], res[0][1]) < 2e-3
        og.G_forward(z, True); og.zero_grads(); gz0 = og.G_backward(gout)
        assert rel(res[1][0], gz0) < 1e-2 and rel(res[1][1], og.grads) < 1e-2
        # D32_st3 (eval mode: no dropout randomness), STNs off the identity
        B = 6
        od = po.Model(po.D32_ST3, Cc, 100, seed=3); p = od.params; p += rng.standard_normal(p.size).astype(np.float32) * 0.01
        d = models.create_D((Cc, 32, 32), True); d.set_params(od.params); d.evaluate()
        x = rng.uniform(0, 1, (B, Cc, 32, 32)).astype(np.float32); go = rng.standard_normal(B).astype(np.float32)
        res = {}
        for be in (0, 1):
            lib.check(L.cg_set_conv_engine(0)); d.forward(x)
            lib.check(L.cg_set_conv_engine(be)); d.zeroGradParameters(); gx = d.backward(x, go)
            res[be] = (gx.copy(), d.get_grads())
        assert rel(res[1][0], res[0][0]) < 2e-3 and rel(res[1][1], res[0][1]) < 2e-3
        od.D_forward(x, None); od.zero_grads(); gx0 = od.D_backward(go)
        assert rel(res[1][0], gx0) < 1e-2 and rel(res[1][1], od.grads) < 1e-2
    finally:
        lib.check(L.cg_set_conv_engine(1))


@pytest.mark.parametrize("Cc,train", [(3, True), (1, True), (3, False)], ids=["rgb-train", "gray-train", "rgb-eval"])
def test_D_forward_backward(Cc, train, engine):
    rng = np.random.default_rng(2)
    B = 6
    od = po.Model(po.D32_ST3, Cc, 100, seed=3)
    p = od.params; p += rng.standard_normal(p.size).astype(np.float32) * 0.01   # STNs off the identity
    d = models.create_D((Cc, 32, 32), True, seed=9)
    assert d.nparams == od.n
    d.set_params(od.params)
    x = rng.uniform(0, 1, (B, Cc, 32, 32)).astype(np.float32)
    d.training() if train else d.evaluate()
    out, pre = d.forward(x, with_pre=True)
    masks = d.get_masks()                       # the masks the GPU drew are replayed in the oracle
    if train:
        assert set(np.unique(masks[:B * 704])) <= {0.0, 1.0} and set(np.unique(masks[B * 704:])) <= {0.0, 2.0}
        keep = masks[:B * 384].mean(); assert 0.7 < keep < 0.9          # SpatialDropout(0.2)
    sig0, pre0 = od.D_forward(x, masks if train else None)
    assert np.abs(pre - pre0).max() < (2e-4 if engine == 0 else 5e-3)
    assert np.abs(out[:, 0] - sig0).max() < 1e-3
    gout = rng.standard_normal(B).astype(np.float32)
    od.zero_grads(); gx0 = od.D_backward(gout)
    d.zeroGradParameters(); gx = d.backward(x, gout)
    assert rel(d.get_grads(), od.grads) < gtol(engine)
    if engine == 0:
        assert rel(gx, gx0) < 1e-2
    else:
        assert l2rel(gx, gx0) < 1e-1      # max-abs is dominated by single max-pool reroutes (statement (c))
    if not train:
        assert np.allclose(masks[:B * 384], 0.8) and np.allclose(masks[-B * 256:], 1.0)


def test_D_masks_can_be_injected():
    rng = np.random.default_rng(4)
    B = 4
    od = po.Model(po.D32_ST3, 3, 100, seed=3)
    d = models.create_D((3, 32, 32), True)
    d.set_params(od.params)
    x = rng.uniform(0, 1, (B, 3, 32, 32)).astype(np.float32)
    masks = po.make_D_masks(B, rng)
    d.set_masks(masks, B, 1)
    out = d.forward(x)
    assert np.array_equal(d.get_masks(), masks)
    sig0, _ = od.D_forward(x, masks)
    assert np.abs(out[:, 0] - sig0).max() < 1e-3


def _closure_inputs(rng, Cc, B):
    real = rng.uniform(0, 1, (B // 2, Cc, 32, 32)).astype(np.float32)
    zD = rng.uniform(-1, 1, (B // 2, 100)).astype(np.float32)
    zG = rng.uniform(-1, 1, (B, 100)).astype(np.float32)
    return real, zD, zG, po.make_D_masks(B, rng), po.make_D_masks(B, rng)


def _gpu_fevalD(L, g, d, cfg, real, zD, maskD, B):
    """adversarial.lua:221-238 + fevalD (:72-112) driven through the per-module C-ABI calls, as the Lua shim would."""
    fake = g.forward(zD)
    inputs = np.concatenate([real, fake]).astype(np.float32)
    targets = np.concatenate([np.ones(B // 2, np.float32), np.zeros(B - B // 2, np.float32)])
    d.zeroGradParameters(); d.set_masks(maskD, B, 1)
    out = d.forward(inputs)[:, 0].copy()
    loss = np.zeros(1, np.float32); df = np.empty(B, np.float32)
    lib.check(L.cg_bce(P(out), P(targets), B, P(loss), P(df)))
    d.backward(inputs, df)
    pen = np.zeros(1, np.float32)
    lib.check(L.cg_penalty_clamp(d.h, cfg.D_L1, cfg.D_L1, cfg.D_L2, cfg.D_clamp, P(pen)))
    return inputs, targets, out, float(loss[0] + pen[0]), d.get_grads()


def _gpu_fevalG(L, g, d, cfg, zG, maskG, B):
    """fevalG_on_D (adversarial.lua:171-215) through the per-module calls."""
    g.zeroGradParameters()
    samples = g.forward(zG)
    d.set_masks(maskG, B, 1)
    out = d.forward(samples)[:, 0].copy()
    loss = np.zeros(1, np.float32); df = np.empty(B, np.float32)
    lib.check(L.cg_bce(P(out), P(np.ones(B, np.float32)), B, P(loss), P(df)))
    gimg = d.backward(samples, df)
    g.backward(zG, gimg)
    pen = np.zeros(1, np.float32)
    lib.check(L.cg_penalty_clamp(g.h, cfg.G_L1, cfg.G_L2, cfg.G_L2, cfg.G_clamp, P(pen)))
    return out, float(loss[0] + pen[0]), gimg, g.get_grads()


@pytest.mark.parametrize("gk,ok,Cc,B", [(lib.G32UPC, po.G32UPC, 3, 8), (lib.G32UP, po.G32UP, 1, 8)], ids=["c2-like", "c1c3-like"])
def test_closures_match_oracle(gk, ok, Cc, B, engine):
    """The two closures of adversarial.train with identical parameters, inputs and dropout masks on both sides.

    What is compared is what is WELL-POSED.  Measured with the oracle alone (profiles/r01_parity_noise_floor.txt):
    two runs of the same fp32 oracle that differ only in summation order agree on a closure's loss to 1e-7 and on
    its gradient to 1e-8..1e-3, but after Adam updates (sign-like steps; max-pool / bilinear-cell routing in D)
    they disagree on 0.36% of G's next steps by more than lr/2 and on the next loss by 8e-4.  So parameters are
    re-synchronised after D's update and gradients, not trajectories, carry the parity claim.
    Gradient bound 1e-2 of max|oracle|: the oracle itself is only 1.3e-3 from a float64 restatement on G32up-c
    (tools/backward_precision_study.py), so nothing tighter can be claimed against it."""
    L = lib.load()
    rng = np.random.default_rng(5)
    og, od = po.Model(ok, Cc, 100, seed=1), po.Model(po.D32_ST3, Cc, 100, seed=2)
    g = models.create_G((Cc, 32, 32), 100, kind=gk); d = models.create_D((Cc, 32, 32), True)
    g.set_params(og.params); g.set_bn_running(og.bn_running); d.set_params(od.params)
    t, ot = adversarial.Trainer(g, d), po.Trainer(og, od)
    cfg, ocfg = lib.default_cfg(B), po.default_cfg(B)
    real, zD, zG, maskD, maskG = _closure_inputs(rng, Cc, B)
    ltol = 1e-4 if engine == 0 else 2e-3
    # ---- fevalD
    inputs, targets, out, f, gD = _gpu_fevalD(L, g, d, cfg, real, zD, maskD, B)
    fake0 = og.G_forward(zD, True)
    assert np.abs(inputs[B // 2:] - fake0).max() < 1e-3                      # fakes fed to D (pixels, north_star bound)
    dout0 = np.zeros(B, np.float32)
    f0 = po.lib().og_fevalD(ot.h, C.byref(ocfg), po.P(np.concatenate([real, fake0]).astype(np.float32)), po.P(targets), po.P(maskD), po.P(dout0))
    assert abs(f - f0) < ltol and np.abs(out - dout0).max() < max(ltol, 1e-3)
    assert rel(gD, od.grads) < gtol(engine)
    # ---- optim.adam on D on both sides, then re-synchronise (see docstring)
    lib.check(L.cg_adam_step(t.h, 0, C.byref(cfg)))
    po.lib().og_adam_step(po.P(od.params), po.P(od.grads), po.P(np.zeros(od.n, np.float32)), po.P(np.zeros(od.n, np.float32)), od.n, 1, 1e-3, 0.9, 0.999, 1e-8)
    pD = d.get_params()
    big = np.abs(od.grads) > 1e-4 * np.abs(od.grads).max()                   # away from sign-noise: both take the same step
    assert np.mean(np.abs(pD - od.params)[big] > 0.5e-3) < 1e-3
    d.set_params(od.params)
    g.set_bn_running(og.bn_running)
    # ---- fevalG_on_D
    out, f, gimg, gG = _gpu_fevalG(L, g, d, cfg, zG, maskG, B)
    f0 = po.lib().og_fevalG_on_D(ot.h, C.byref(ocfg), po.P(zG), po.P(maskG))
    assert abs(f - f0) < ltol
    assert rel(gG, og.grads) < gtol(engine)


def test_fused_step_equals_unfused_sequence():
    """cg_train_step (one C call per step) must do exactly what the per-module sequence does: GPU against GPU, same
    kernels, so only the fp32 atomics of the bilinear scatter-add may differ."""
    L = lib.load()
    rng = np.random.default_rng(8)
    Cc, B = 3, 8
    og, od = po.Model(po.G32UPC, Cc, 100, seed=1), po.Model(po.D32_ST3, Cc, 100, seed=2)
    seedp, seedd = og.params.copy(), od.params.copy()
    assert np.isfinite(seedp).all() and np.isfinite(seedd).all()
    real, zD, zG, maskD, maskG = _closure_inputs(rng, Cc, B)
    cfg = lib.default_cfg(B)
    ga = models.create_G((Cc, 32, 32), 100); da = models.create_D((Cc, 32, 32), True); ga.set_params(seedp); da.set_params(seedd)
    ta = adversarial.Trainer(ga, da)
    da.set_masks(np.stack([maskD, maskG]), B, 2)
    lD, lG, dout = ta.step(cfg, real[None], zD[None], zG[None])
    gb = models.create_G((Cc, 32, 32), 100); db = models.create_D((Cc, 32, 32), True); gb.set_params(seedp); db.set_params(seedd)
    tb = adversarial.Trainer(gb, db)
    _, _, outb, fD, _ = _gpu_fevalD(L, gb, db, cfg, real, zD, maskD, B)
    lib.check(L.cg_adam_step(tb.h, 0, C.byref(cfg)))
    _, fG, _, gGb = _gpu_fevalG(L, gb, db, cfg, zG, maskG, B)
    lib.check(L.cg_adam_step(tb.h, 1, C.byref(cfg)))
    assert abs(lD[0] - fD) < 1e-5 and abs(lG[0] - fG) < 1e-4 and np.abs(dout - outb).max() < 1e-5
    assert rel(ga.get_grads(), gGb) < 1e-3
    assert np.mean(np.abs(ga.get_params() - gb.get_params()) > 0.5e-3) < 1e-3
    assert np.abs(da.get_params() - db.get_params()).max() < 1e-5


def test_graph_replay_matches_eager():
    """cg_train_step replays the step as a CUDA graph from the third call of a configuration on.  Same initial state, same
    inputs, graphs off vs on: the first replayed step (call 3) must reproduce the eager loss and D outputs; afterwards only the
    bilinear scatter's atomic order differs and the trajectories may drift at the floor measured for the oracle itself (8e-4 on
    the loss).  Adam's step count and the dropout RNG offset are device-side, so replay advances both (losses keep changing)."""
    L = lib.load()
    Cc, B = 3, 8
    rng = np.random.default_rng(11)
    steps = [(_closure_inputs(rng, Cc, B)) for _ in range(6)]
    out = {}
    try:
        for mode in (0, 1):
            lib.check(L.cg_set_graph_mode(mode))
            g = models.create_G((Cc, 32, 32), 100, seed=1); d = models.create_D((Cc, 32, 32), True, seed=2)
            t = adversarial.Trainer(g, d)
            L.cg_reset_launch_count()
            rec = []
            for real, zD, zG, _, _ in steps:
                n0 = L.cg_launch_count()
                lD, lG, dout = t.step(lib.default_cfg(B), real[None], zD[None], zG[None])
                rec.append((float(lD[0]), float(lG[0]), dout.copy(), L.cg_launch_count() - n0))
            out[mode] = (rec, g.get_params(), d.get_params())
    finally:
        lib.check(L.cg_set_graph_mode(1))
    e, r = out[0][0], out[1][0]
    for i in range(6):
        # lossD and D's outputs are computed BEFORE any update inside the step: tight through the first replay (call 3).
        # lossG is computed after D's Adam update in the same step, i.e. behind sign-like steps that amplify the atomic-order
        # noise of the bilinear scatter: it gets the trajectory floor (measured on B200 at the first replay: lossD 3e-6, lossG 3.6e-4).
        tolD = 1e-4 if i < 2 else 5e-3   # from the third call on both runs are two Adam updates past their common start: the trajectory floor (measured 3e-6 .. 1.4e-4 at i = 2)
        tolG = 5e-3 if i < 3 else 2e-2   # two trajectories four or five Adam updates apart at batch 8 (measured up to 7.6e-3 at i = 4: the scatter-add's atomic order, amplified by sign-like Adam steps)
        assert abs(e[i][0] - r[i][0]) < tolD and abs(e[i][1] - r[i][1]) < tolG, (i, e[i][:2], r[i][:2])
        assert np.abs(e[i][2] - r[i][2]).max() < max(tolD, 1e-3)
        assert r[i][3] == e[i][3] > 300, "replay accounts for the same number of kernel launches as the eager step"
    assert len({round(x[0], 6) for x in r}) == 6, "every replay sees new dropout masks and a new Adam step (losses differ)"
    assert np.mean(np.abs(out[0][1] - out[1][1]) > 0.5e-3) < 2e-2 and np.mean(np.abs(out[0][2] - out[1][2]) > 0.5e-3) < 2e-2


def test_sampler_path_matches_oracle():
    """sample.lua's path -- G and D in evaluate() mode (sample.lua:211,216), images created in chunks of
    OPT.batchSize with a ragged tail (nn_utils.lua:45-69), D's predictions sorted (nn_utils.lua:89-117).  The oracle runs the
    same chunks; G's running statistics are first moved off their initial values by two training-mode forwards on both sides."""
    from catgen import nn_utils
    Cc, N, bs = 3, 20, 8
    rng = np.random.default_rng(41)
    og, od = po.Model(po.G32UPC, Cc, 100, seed=5), po.Model(po.D32_ST3, Cc, 100, seed=6)
    g = models.create_G((Cc, 32, 32), 100); d = models.create_D((Cc, 32, 32), True)
    g.set_params(og.params); d.set_params(od.params)
    for _ in range(2):
        zw = nn_utils.createNoiseInputs(16, 100, rng)
        g.forward(zw); og.G_forward(zw, train=True)
    assert rel(g.get_bn_running(), og.bn_running) < 5e-3
    nn_utils.switchToEvaluationMode(g, d)
    z = nn_utils.createNoiseInputs(N, 100, rng)
    images = nn_utils.createImagesFromNoise(g, z, bs)
    assert images.shape == (N, Cc, 32, 32) and len(nn_utils.createImagesFromNoise(g, z, bs, outputAsList=True)) == N
    ref = np.concatenate([og.G_forward(z[s:s + bs], train=False) for s in range(0, N, bs)])
    assert np.abs(images - ref).max() < 1e-3                      # north_star pixel tolerance
    assert rel(g.get_bn_running(), og.bn_running) < 5e-3            # evaluate() must not touch the running statistics
    best, preds = nn_utils.sortImagesByPrediction(d, images, False, 6, bs)
    worst, wpreds = nn_utils.sortImagesByPrediction(d, images, True, 64, bs)
    ref_pred = np.concatenate([od.D_forward(ref[s:s + bs], None)[0] for s in range(0, N, bs)])
    assert len(best) == 6 and len(worst) == N
    assert np.abs(np.array(preds) - np.sort(ref_pred)[::-1][:6]).max() < 1e-3
    assert np.abs(np.array(wpreds) - np.sort(ref_pred)).max() < 1e-3
    assert all(a >= b for a, b in zip(preds, preds[1:])) and all(a <= b for a, b in zip(wpreds, wpreds[1:]))
    top = int(np.argmax(np.concatenate([d.forward(images[s0:s0 + bs])[:, 0] for s0 in range(0, N, bs)])))
    assert np.array_equal(best[0], images[top])
    nn_utils.switchToTrainingMode(g, d)


def test_concurrent_streams_match_single_stream():
    """The step issues independent work on concurrent streams (D's four branches, each layer's weight-gradient chain, fevalG's
    generator forward; cg_set_concurrency).  Whatever the interleaving, results must equal the single-stream program order:
    same kernels on the same operands, so only the bilinear scatter's atomic order may differ.  Repeated to give a race a
    chance to show; B = 16 keeps several CTAs per kernel in flight."""
    L = lib.load()
    Cc, B = 3, 16
    rng = np.random.default_rng(31)
    x = rng.uniform(0, 1, (B, Cc, 32, 32)).astype(np.float32)
    z = rng.uniform(-1, 1, (B, 100)).astype(np.float32)
    go = rng.standard_normal(B).astype(np.float32)
    gimg = rng.standard_normal((B, Cc, 32, 32)).astype(np.float32)
    masks = po.make_D_masks(B, rng)
    ref = None
    try:
        for rep, mode in enumerate((0, 1, 1, 1)):
            lib.check(L.cg_set_concurrency(mode))
            d = models.create_D((Cc, 32, 32), True, seed=2); g = models.create_G((Cc, 32, 32), 100, seed=1)
            d.training(); d.set_masks(masks, B, 1)
            out = d.forward(x).copy(); d.zeroGradParameters(); gx = d.backward(x, go).copy(); gD = d.get_grads()
            img = g.forward(z).copy(); g.zeroGradParameters(); gz = g.backward(z, gimg).copy(); gG = g.get_grads()
            cur = (out, gx, gD, img, gz, gG)
            if ref is None: ref = cur
            else:
                for name, a, b in zip(("D out", "D gradInput", "D grads", "G out", "G gradInput", "G grads"), cur, ref):
                    assert rel(a, b) < 2e-5, (rep, name, rel(a, b))
        # and the fused step: three steps from the same state, eager, with and without concurrency
        steps = [(_closure_inputs(rng, Cc, 8)) for _ in range(3)]
        rec = {}
        lib.check(L.cg_set_graph_mode(0))
        for mode in (0, 1):
            lib.check(L.cg_set_concurrency(mode))
            g = models.create_G((Cc, 32, 32), 100, seed=1); d = models.create_D((Cc, 32, 32), True, seed=2)
            t = adversarial.Trainer(g, d)
            ls = [t.step(lib.default_cfg(8), real[None], zD[None], zG[None]) for real, zD, zG, _, _ in steps]
            rec[mode] = ([(float(a[0]), float(b[0])) for a, b, _ in ls], g.get_params(), d.get_params())
        assert abs(rec[0][0][0][0] - rec[1][0][0][0]) < 1e-5 and abs(rec[0][0][0][1] - rec[1][0][0][1]) < 1e-4
        for i in range(3):
            assert abs(rec[0][0][i][0] - rec[1][0][i][0]) < 5e-3 and abs(rec[0][0][i][1] - rec[1][0][i][1]) < 5e-3
        assert np.mean(np.abs(rec[0][1] - rec[1][1]) > 0.5e-3) < 2e-2 and np.mean(np.abs(rec[0][2] - rec[1][2]) > 0.5e-3) < 2e-2
    finally:
        lib.check(L.cg_set_concurrency(1)); lib.check(L.cg_set_graph_mode(1))


def test_dead_grad_elimination_is_unobservable():
    """Inside fevalG the reference's MODEL_D:backward also accumulates D's parameter gradients, which the next fevalD zeroes
    unread (adversarial.lua:193 vs :78).  cg_train_step skips that accumulation by default (cg_set_dead_grad_elim).  The input-
    gradient path is the same kernels on the same operands either way, so losses, D outputs and both parameter vectors must
    agree to the bilinear scatter's atomic-order floor -- and the switch must actually remove launches."""
    L = lib.load()
    Cc, B = 3, 8
    rng = np.random.default_rng(21)
    steps = [(_closure_inputs(rng, Cc, B)) for _ in range(3)]
    out = {}
    try:
        lib.check(L.cg_set_graph_mode(0))
        for mode in (0, 1):
            lib.check(L.cg_set_dead_grad_elim(mode))
            g = models.create_G((Cc, 32, 32), 100, seed=1); d = models.create_D((Cc, 32, 32), True, seed=2)
            t = adversarial.Trainer(g, d)
            rec = []
            for real, zD, zG, _, _ in steps:
                n0 = L.cg_launch_count()
                lD, lG, dout = t.step(lib.default_cfg(B), real[None], zD[None], zG[None])
                rec.append((float(lD[0]), float(lG[0]), dout.copy(), L.cg_launch_count() - n0))
            out[mode] = (rec, g.get_params(), d.get_params(), d.get_grads())
    finally:
        lib.check(L.cg_set_graph_mode(1)); lib.check(L.cg_set_dead_grad_elim(1))
    keep, skip = out[0], out[1]
    assert abs(keep[0][0][0] - skip[0][0][0]) < 1e-5 and abs(keep[0][0][1] - skip[0][0][1]) < 1e-4      # first step: before any drift
    assert np.abs(keep[0][0][2] - skip[0][0][2]).max() < 1e-5
    for i in range(3):
        assert abs(keep[0][i][0] - skip[0][i][0]) < 5e-3 and abs(keep[0][i][1] - skip[0][i][1]) < 5e-3
        assert skip[0][i][3] < keep[0][i][3] - 50, "the switch removes D's weight-gradient launches from fevalG"
    assert np.mean(np.abs(keep[1] - skip[1]) > 0.5e-3) < 2e-2 and np.mean(np.abs(keep[2] - skip[2]) > 0.5e-3) < 2e-2
    assert rel(keep[3], skip[3]) > 1e-2, "with the accumulation kept, D's gradient vector holds fevalG's (dead) gradients instead"


def test_fused_step_losses_track_oracle(engine):
    """Two fused steps against the oracle's og_train_step: the first D update is exact-input parity (tight), what
    follows it has passed through Adam and is compared at the measured trajectory floor (oracle vs itself: 8e-4)."""
    rng = np.random.default_rng(5)
    Cc, B = 3, 8
    og, od = po.Model(po.G32UPC, Cc, 100, seed=1), po.Model(po.D32_ST3, Cc, 100, seed=2)
    g = models.create_G((Cc, 32, 32), 100); d = models.create_D((Cc, 32, 32), True)
    g.set_params(og.params); g.set_bn_running(og.bn_running); d.set_params(od.params)
    t, ot = adversarial.Trainer(g, d), po.Trainer(og, od)
    for step in range(2):
        real, zD, zG, mD, mG = _closure_inputs(rng, Cc, B)
        masks = np.stack([mD, mG]); d.set_masks(masks, B, 2)
        lD, lG, dout = t.step(lib.default_cfg(B), real[None], zD[None], zG[None])
        lD0, lG0, dout0 = ot.step(po.default_cfg(B), real[None], zD[None], zG[None], masks)
        tol = (2e-4 if engine == 0 else 2e-3) if step == 0 else 5e-3
        assert abs(lD[0] - lD0[0]) < tol and np.abs(dout - dout0).max() < max(tol, 1e-3)
        assert abs(lG[0] - lG0[0]) < 5e-3


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_properties_c2():
    """BASELINE config c2 (G32up-c, RGB, B=128): properties that need no oracle run at that size."""
    rng = np.random.default_rng(6)
    B = 128
    g = models.create_G((3, 32, 32), 100); d = models.create_D((3, 32, 32), True)
    z = rng.uniform(-1, 1, (B, 100)).astype(np.float32)
    px = g.forward(z)
    assert px.shape == (B, 3, 32, 32) and np.isfinite(px).all() and px.min() > 0 and px.max() < 1   # sigmoid range
    px2 = g.forward(z)
    assert np.array_equal(px, px2), "G forward is deterministic (split-K partials are reduced in fixed order)"
    # batch statistics: permuting the batch permutes the output (BN couples samples only through symmetric sums)
    perm = rng.permutation(B)
    # the persistent conv kernel starts each tile's K loop at a slice that depends on the tile's position in the launch, so permuting the
    # batch changes the fp32 summation order: accumulation-order noise through four conv + BN layers (north_star tolerance: 1e-3)
    assert np.abs(g.forward(z[perm]) - px[perm]).max() < 5e-4
    d.evaluate()
    o1 = d.forward(px); o2 = d.forward(px[perm])
    assert np.abs(o1[perm] - o2).max() < 1e-5, "D is per-sample in eval mode"
    assert ((o1 > 0) & (o1 < 1)).all()
    t = adversarial.Trainer(g, d); d.training()
    cfg = lib.default_cfg(B)
    real = rng.uniform(0, 1, (1, B // 2, 3, 32, 32)).astype(np.float32)
    pD0, pG0 = d.get_params(), g.get_params()
    lD, lG, dout = t.step(cfg, real, rng.uniform(-1, 1, (1, B // 2, 100)).astype(np.float32), rng.uniform(-1, 1, (1, B, 100)).astype(np.float32))
    assert np.isfinite(lD).all() and np.isfinite(lG).all() and dout.shape == (B,)
    dD, dG = d.get_params() - pD0, g.get_params() - pG0
    # First Adam step in closed form (t = 1): m = (1-b1) g, v = (1-b2) g^2, so
    #   |dp| = lr * |g| / (|g| + eps / sqrt(1 - b2)),   eps / sqrt(1 - b2) = 3.16e-7,
    # for EVERY element, whatever the gradient magnitude (measured on B200: G's gradients at init are ~1e-7, so
    # most steps are well below lr -- an earlier version of this test wrongly assumed most would equal lr).
    # After the step the library's flat gradient still holds the clamped gradient Adam consumed (G only: D's
    # buffer also receives the G-phase accumulation, adversarial.lua:192).
    assert np.abs(dD).max() <= 1.001e-3 and np.abs(dG).max() <= 1.001e-3
    gG = g.get_grads().astype(np.float64)
    expect = -1e-3 * gG / (np.abs(gG) + 1e-8 / np.sqrt(1 - 0.999))
    assert np.abs(gG).max() > 0
    # parameters are O(0.1) in fp32: the update itself is only resolved to ~1e-8 absolute
    assert np.abs(dG - expect).max() < 3e-8 + 1e-3 * 2e-3, np.abs(dG - expect).max()
