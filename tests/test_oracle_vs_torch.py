"""Pins the C oracle against an independent PyTorch-CPU autograd restatement.
The reference ships no golden vectors (parity unpinned), so this is the strongest check available."""
import numpy as np
import pytest
import torch

from oracle import pyoracle as po
from oracle import torch_ref as tr

torch.set_num_threads(8)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def _truth_and_fp32(fn, flat, inp, gout):
    """Run the torch restatement in float64 (truth) and float32; return {dtype: (out, ginp, gflat)}."""
    res = {}
    for dt in (torch.float64, torch.float32):
        f = torch.tensor(flat).to(dt).requires_grad_()
        x = torch.tensor(inp).to(dt).requires_grad_()
        o = fn(f, x)
        o.backward(torch.tensor(gout).to(dt))
        res[dt] = (o.detach().numpy(), x.grad.numpy(), f.grad.numpy())
    return res


def _check_grad(mine, truth, fp32, floor):
    """`mine` must be as close to the float64 truth as an independent fp32 implementation is.
    Tolerance = 3x the fp32 party's own error + `floor` (relative to max|truth|): a semantic error shows
    up as mine >> fp32-noise, while honest roundoff does not.  Measured on this box (B=4, seed 1):
    G32up-c oracle 3.7e-4 / torch-fp32 6.0e-4 (ill-conditioned tiny-batch BN backward at 8x8);
    G32up 2.9e-6 / 7.6e-7."""
    e_mine, e_fp32 = rel(mine, truth), rel(fp32, truth)
    assert e_mine <= 3.0 * e_fp32 + floor, (e_mine, e_fp32)


@pytest.mark.parametrize("kind,name,C", [(po.G32UPC, "G32UPC", 3), (po.G32UP, "G32UP", 1), (po.G32UP, "G32UP", 3)])
def test_G_fwd_bwd(kind, name, C):
    rng = np.random.default_rng(1)
    B = 4
    g = po.Model(kind, C, 100, seed=1)
    z = rng.uniform(-1, 1, (B, 100)).astype(np.float32)
    gout = rng.standard_normal((B, C, 32, 32)).astype(np.float32)
    out = g.G_forward(z, train=True)
    g.zero_grads()
    gz = g.G_backward(gout)
    r = _truth_and_fp32(lambda f, x: tr.G_forward(f, x, name, C), g.params.copy(), z, gout)
    o64, gz64, gp64 = r[torch.float64]
    _, gz32, gp32 = r[torch.float32]
    assert np.abs(out - o64).max() < 1e-5          # sigmoid pixels; measured 6.7e-7
    _check_grad(gz, gz64, gz32, 2e-5)
    _check_grad(g.grads, gp64, gp32, 2e-5)


@pytest.mark.parametrize("C,train", [(3, True), (1, True), (3, False)])
def test_D_fwd_bwd(C, train):
    rng = np.random.default_rng(2)
    B = 4
    d = po.Model(po.D32_ST3, C, 100, seed=3)
    # move the STNs off the identity so rotation/scale/translation gradients are exercised
    p = d.params
    p += rng.standard_normal(p.size).astype(np.float32) * 0.01
    x = rng.uniform(0, 1, (B, C, 32, 32)).astype(np.float32)
    masks = po.make_D_masks(B, rng) if train else None
    sig, pre = d.D_forward(x, masks)
    gout = rng.standard_normal(B).astype(np.float32)
    d.zero_grads()
    gx = d.D_backward(gout)
    mt = None if masks is None else masks
    def fn(f, xx):
        m = None if mt is None else torch.tensor(mt).to(f.dtype)
        return tr.D_forward(f, xx, m, C)[0]
    r = _truth_and_fp32(fn, d.params.copy(), x, gout)
    s64, gx64, gp64 = r[torch.float64]
    _, gx32, gp32 = r[torch.float32]
    pre64 = tr.D_forward(torch.tensor(d.params.copy()).double(), torch.tensor(x).double(),
                         None if mt is None else torch.tensor(mt).double(), C)[1].numpy()
    assert np.abs(pre - pre64).max() < 5e-5
    assert np.abs(sig - s64).max() < 2e-5
    _check_grad(gx, gx64, gx32, 2e-5)
    _check_grad(d.grads, gp64, gp32, 2e-5)


def test_bce_matches_formula():
    L = po.lib()
    p = np.array([0.1, 0.9, 0.5, 1e-7, 1 - 1e-7], np.float32)
    t = np.array([0, 1, 1, 0, 1], np.float32)
    f = L.og_bce_fwd(po.P(p), po.P(t), 5)
    g = np.empty(5, np.float32)
    L.og_bce_bwd(po.P(p), po.P(t), po.P(g), 5)
    pt = torch.tensor(p.astype(np.float64), requires_grad=True)
    l = tr.bce(pt, torch.tensor(t.astype(np.float64)))
    l.backward()
    assert abs(f - l.item()) < 1e-6
    assert rel(g, pt.grad.numpy()) < 1e-5


def test_adam_matches_torch7_formula():
    # optim.adam: eps added to sqrt(v) BEFORE bias correction
    L = po.lib()
    rng = np.random.default_rng(0)
    n = 1000
    x = rng.standard_normal(n).astype(np.float32)
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    x64, m64, v64 = x.astype(np.float64), np.zeros(n), np.zeros(n)
    for t in range(1, 4):
        g = rng.standard_normal(n).astype(np.float32)
        L.og_adam_step(po.P(x), po.P(g), po.P(m), po.P(v), n, t, 1e-3, 0.9, 0.999, 1e-8)
        m64 = 0.9 * m64 + 0.1 * g
        v64 = 0.999 * v64 + 0.001 * g.astype(np.float64) ** 2
        x64 -= 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m64 / (np.sqrt(v64) + 1e-8)
    assert np.abs(x - x64).max() < 1e-6


def test_leakyrelu_grad_at_zero_is_one():
    # LeakyReLU.lua:21-31: sign(negative)+1 == 1 at x == 0
    L = po.lib()
    x = np.array([-1.0, 0.0, 2.0], np.float32)
    gy = np.ones(3, np.float32)
    gx = np.empty(3, np.float32)
    L.og_leakyrelu_bwd(po.P(x), po.P(gy), 0.333, po.P(gx), 3)
    assert np.allclose(gx, [0.333, 1.0, 1.0])


def test_param_counts():
    # SURVEY.md section 8a per-layer tables
    assert po.Model(po.G32UPC, 3).n == 5191687
    assert po.Model(po.G32UPC, 1).n == 5189381
    assert po.Model(po.G32UP, 3).n == 2470406
    assert po.Model(po.D32_ST3, 3).n == 6664777


@pytest.mark.parametrize("C", [1, 3])
def test_V32_forward_matches_torch(C):
    """F3: the oracle's create_V32 composition (evaluate() mode) against an independent torch.nn.functional restatement."""
    flat, run = po.V32_init(C, seed=11)
    assert flat.size == po.V32_nparams(C) == {1: 6285954, 3: 6288258}[C]
    rng = np.random.default_rng(4)
    x = rng.uniform(0, 1, (6, C, 32, 32)).astype(np.float32)
    got = po.V_forward(flat, run, x)
    want = tr.V_forward(torch.from_numpy(flat).double(), torch.from_numpy(run).double(), torch.from_numpy(x).double(), C).numpy()
    assert got.shape == (6, 2)
    np.testing.assert_allclose(got.sum(1), 1.0, atol=1e-6)
    assert np.abs(got - want).max() < 2e-5
    assert abs(po.rateWithV(flat, run, x) - (1 - want[:, 0].mean())) < 2e-5
