"""The PyTorch-CPU restatement of the training step (oracle/torch_step.py, the "B-mkl" CPU baseline of BASELINE.md section 3) against
the C oracle's og_train_step on identical inputs and dropout masks: same losses, same D outputs, and the same first Adam step wherever
the gradient is not sign-noise.  Two independently written restatements of adversarial.lua:221-266 agreeing is also what pins the
STEP ORDER (separate G forward for the fakes, penalty before clamp, D updated before the G phase)."""
import numpy as np
import pytest

from oracle import pyoracle as po, torch_step as ts


@pytest.mark.parametrize("kind,okind,C,d_it", [("G32UPC", po.G32UPC, 3, 1), ("G32UP", po.G32UP, 1, 2)], ids=["c2-like", "c1c3-like-d2"])
def test_torch_step_matches_oracle_step(kind, okind, C, d_it):
    rng = np.random.default_rng(0)
    B = 8
    og, od = po.Model(okind, C, 100, seed=1), po.Model(po.D32_ST3, C, 100, seed=2)
    G, D = ts.Net(np.array(og.params)), ts.Net(np.array(od.params))
    cfg = po.default_cfg(B, d_it, 1)
    real = rng.uniform(0, 1, (d_it, B // 2, C, 32, 32)).astype(np.float32)
    zD = rng.uniform(-1, 1, (d_it, B // 2, 100)).astype(np.float32)
    zG = rng.uniform(-1, 1, (1, B, 100)).astype(np.float32)
    masks = np.stack([po.make_D_masks(B, rng) for _ in range(d_it + 1)])
    lD0, lG0, d0 = po.Trainer(og, od).step(cfg, real, zD, zG, masks)
    lD, lG, d = ts.train_step(G, D, cfg, real, zD, zG, masks, kind, C)
    assert abs(lD[0] - lD0[0]) < 1e-5
    for k in range(1, d_it):
        assert abs(lD[k] - lD0[k]) < 5e-3          # behind an Adam update: trajectory floor (profiles/r01_parity_noise_floor.txt)
    assert abs(lG[0] - lG0[0]) < 5e-3
    assert np.abs(d - d0).max() < (1e-5 if d_it == 1 else 1e-2)
    # Adam's first step is +-lr wherever |g| >> eps: the two implementations may only disagree where the gradient is sign-noise
    pd, pg = D.p.detach().numpy(), G.p.detach().numpy()
    assert np.mean(np.abs(pd - od.params) > 0.5e-3) < 2e-2 and np.mean(np.abs(pg - og.params) > 0.5e-3) < 2e-2
    assert D.t == d_it and G.t == 1
