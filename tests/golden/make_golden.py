"""Generates tests/golden/*.npz -- float64 reference values for small seeded cases of the hot path.

Where the numbers come from: the reference (Lua/Torch7) cannot run in this environment and ships no golden vectors, so these
are produced by the float64 PyTorch-autograd restatement of the reference's networks in oracle/torch_ref.py (models.lua:138-160,
196-228, 640-711, 814-906), which shares no code with the C oracle or the CUDA kernels.  They pin BOTH: tests/test_golden.py
checks the oracle against them on CPU and the CUDA path against them on the GPU box, where neither this script nor torch_ref runs.

Parameters are not stored (20 MB per network): they are regenerated from the oracle's seeded initialiser and guarded by a
checksum stored in the fixture.  Parameter gradients are stored as a fixed random sample of entries plus norms.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz (deterministic: seeds only)"""
import os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from oracle import pyoracle as po
from oracle import torch_ref as tr

torch.set_num_threads(8)
NSAMP = 4096

CASES = [   # name, network, torch_ref kind, C, B, params seed, data seed, D in training mode
    ("G32up_y_b4", po.G32UP, "G32UP", 1, 4, 11, 101, None),
    ("G32upc_rgb_b4", po.G32UPC, "G32UPC", 3, 4, 12, 102, None),
    ("G32upc_rgb_b6", po.G32UPC, "G32UPC", 3, 6, 13, 103, None),
    ("D32st3_rgb_b4_train", po.D32_ST3, None, 3, 4, 14, 104, True),
    ("D32st3_y_b4_train", po.D32_ST3, None, 1, 4, 15, 105, True),
    ("D32st3_rgb_b5_eval", po.D32_ST3, None, 3, 5, 16, 106, False),
]


def params_for(net, C, seed, data_seed):
    m = po.Model(net, C, 100, seed=seed)
    p = m.params.copy()
    if net == po.D32_ST3:   # move the transformers off the identity so rotation/scale/translation gradients are exercised
        p += np.random.default_rng(data_seed + 1000).standard_normal(p.size).astype(np.float32) * 0.01
    return p


def checksum(p):
    p = p.astype(np.float64)
    return np.array([p.sum(), (p * p).sum(), p[::997].sum()])


def run(fn, flat, inp, gout, dt):
    f = torch.tensor(flat).to(dt).requires_grad_()
    x = torch.tensor(inp).to(dt).requires_grad_()
    o = fn(f, x)
    o.backward(torch.tensor(gout).to(dt))
    return o.detach().numpy().astype(np.float64), x.grad.numpy().astype(np.float64), f.grad.numpy().astype(np.float64)


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def main():
    for name, net, kind, C, B, seed, dseed, train in CASES:
        rng = np.random.default_rng(dseed)
        p = params_for(net, C, seed, dseed)
        fx = {"C": C, "B": B, "seed": seed, "data_seed": dseed, "params_checksum": checksum(p)}
        if net != po.D32_ST3:
            inp = rng.uniform(-1, 1, (B, 100)).astype(np.float32)
            gout = rng.standard_normal((B, C, 32, 32)).astype(np.float32)
            fn = lambda f, x: tr.G_forward(f, x, kind, C)
        else:
            inp = rng.uniform(0, 1, (B, C, 32, 32)).astype(np.float32)
            masks = po.make_D_masks(B, rng) if train else None
            gout = rng.standard_normal(B).astype(np.float32)
            fx["masks"] = masks if masks is not None else np.zeros(0, np.float32)
            fx["train"] = int(bool(train))
            fn = lambda f, x: tr.D_forward(f, x, None if masks is None else torch.tensor(masks).to(f.dtype), C)[0]
        o64, gi64, gp64 = run(fn, p, inp, gout, torch.float64)
        o32, gi32, gp32 = run(fn, p, inp, gout, torch.float32)
        idx = np.sort(np.random.default_rng(dseed + 7).choice(p.size, NSAMP, replace=False))
        fx.update(inp=inp, gout=gout, out=o64, ginp=gi64, gparam_idx=idx, gparam_sample=gp64[idx],
                  gparam_l2=np.sqrt((gp64 * gp64).sum()), gparam_max=np.abs(gp64).max(),
                  # how far an independent fp32 implementation lands from the float64 truth on this very case
                  fp32_err=np.array([np.abs(o32 - o64).max(), rel(gi32, gi64), rel(gp32[idx], gp64[idx]) * np.abs(gp64[idx]).max() / np.abs(gp64).max()]))
        if net == po.D32_ST3:
            fx["pre"] = tr.D_forward(torch.tensor(p).double(), torch.tensor(inp).double(),
                                     None if masks is None else torch.tensor(masks).double(), C)[1].numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
        print("%-24s out %s  |ginp|max %.3e  |gparam|max %.3e  fp32 party: out %.1e ginp %.1e gparam %.1e" %
              (name, o64.shape, np.abs(gi64).max(), fx["gparam_max"], *fx["fp32_err"]))


if __name__ == "__main__":
    main()
