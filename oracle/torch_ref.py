"""PyTorch-CPU restatement of G32up / G32up-c / D32_st3 used ONLY to cross-check the C oracle
(SURVEY.md A.10).  It is never on a product or timed path.  Parameters come from the same flat
vector (nn getParameters() order) the oracle and the CUDA library use.
"""
import math

import torch
import torch.nn.functional as F


class Cursor:
    def __init__(self, flat):
        self.flat, self.o = flat, 0

    def take(self, *shape):
        n = math.prod(shape)
        t = self.flat[self.o:self.o + n].view(*shape)
        self.o += n
        return t


def g_spec(kind, C):
    if kind == "G32UPC":   # models.lua:196-228
        return 512, 4, [(1, 512, 512, 3, 1), (1, 512, 256, 3, 1), (1, 256, 128, 5, 1), (0, 128, C, 3, 0)]
    return 128, 8, [(1, 128, 256, 5, 1), (1, 256, 128, 5, 1), (0, 128, C, 3, 0)]   # models.lua:138-160


def G_forward(flat, z, kind="G32UPC", C=3, nz=100):
    C0, s0, stages = g_spec(kind, C)
    c = Cursor(flat)
    W, b, pw = c.take(C0 * s0 * s0, nz), c.take(C0 * s0 * s0), c.take(1)
    x = F.prelu(F.linear(z, W, b), pw).view(-1, C0, s0, s0)
    for up, Ci, Co, k, bn in stages:
        if up:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        W, b = c.take(Co, Ci, k, k), c.take(Co)
        x = F.conv2d(x, W, b, padding=(k - 1) // 2)
        if bn:
            g, bt, pw = c.take(Co), c.take(Co), c.take(1)
            x = F.prelu(F.batch_norm(x, None, None, g, bt, training=True, momentum=0.1, eps=1e-5), pw)
        else:
            x = torch.sigmoid(x)
    assert c.o == flat.numel()
    return x


def leaky(x, s=0.333):
    return torch.where(x >= 0, x, s * x)


def affine_matrix(theta, rot, scl, trn):
    B = theta.shape[0]
    M = torch.eye(3, dtype=theta.dtype).expand(B, 3, 3)
    p = 0
    if rot:
        a = theta[:, p]; p += 1
        R = torch.zeros(B, 3, 3, dtype=theta.dtype)
        R[:, 0, 0] = torch.cos(a); R[:, 0, 1] = -torch.sin(a)
        R[:, 1, 0] = torch.sin(a); R[:, 1, 1] = torch.cos(a); R[:, 2, 2] = 1
        M = M @ R
    if scl:
        s = theta[:, p]; p += 1
        S = torch.zeros(B, 3, 3, dtype=theta.dtype)
        S[:, 0, 0] = s; S[:, 1, 1] = s; S[:, 2, 2] = 1
        M = M @ S
    if trn:
        T = torch.eye(3, dtype=theta.dtype).repeat(B, 1, 1)
        T[:, 0, 2] = theta[:, p]; T[:, 1, 2] = theta[:, p + 1]
        M = M @ T
    return M[:, :2, :]


def stn(c, x, ch, S, rot, scl, trn):
    """models.lua:814-906.  stn grid channel 0 = y, 1 = x; F.grid_sample wants (x, y)."""
    nth = int(rot) + int(scl) + 2 * int(trn)
    f = 16 * (S // 4) ** 2
    c1W, c1b, c2W, c2b = c.take(16, ch, 3, 3), c.take(16), c.take(16, 16, 3, 3), c.take(16)
    l1W, l1b, l2W, l2b = c.take(64, f), c.take(64), c.take(nth, 64), c.take(nth)
    h = F.avg_pool2d(x, 2)
    h = leaky(F.conv2d(h, c1W, c1b, padding=1))
    h = leaky(F.conv2d(h, c2W, c2b, padding=1))
    h = F.avg_pool2d(h, 2).reshape(-1, f)
    theta = F.linear(leaky(F.linear(h, l1W, l1b)), l2W, l2b)
    A = affine_matrix(theta, rot, scl, trn)                      # [B,2,3] acting on (y, x, 1)
    lin = torch.linspace(-1, 1, S, dtype=x.dtype)
    yy, xx = torch.meshgrid(lin, lin, indexing="ij")
    base = torch.stack([yy, xx, torch.ones_like(yy)], -1).view(1, S * S, 3)
    grid_yx = (base @ A.transpose(1, 2)).view(-1, S, S, 2)
    grid_xy = torch.stack([grid_yx[..., 1], grid_yx[..., 0]], -1)
    return F.grid_sample(x, grid_xy, mode="bilinear", padding_mode="zeros", align_corners=True)


def D_forward(flat, x, masks=None, C=3):
    """models.lua:640-711.  masks: flat multipliers in oracle layout, or None for eval mode."""
    B = x.shape[0]
    if masks is None:
        masks = torch.cat([torch.full((B * (64 * 4 + 128),), 0.8), torch.full((B * 320,), 0.5),
                           torch.ones(B * 256)]).to(x.dtype)
    mo = [0]

    def mk(n):
        t = masks[mo[0]:mo[0] + n]; mo[0] += n
        return t

    c = Cursor(flat)
    h = stn(c, x, C, 32, True, False, False)
    W, b, pw = c.take(64, C, 3, 3), c.take(64), c.take(1)
    h = F.prelu(F.conv2d(h, W, b, padding=1), pw)
    W, b, pw = c.take(64, 64, 3, 3), c.take(64), c.take(1)
    h = F.prelu(F.conv2d(h, W, b, padding=1), pw)
    T = F.avg_pool2d(h, 2) * mk(B * 64).view(B, 64, 1, 1)
    outs = []
    for br in range(4):
        Co, k1, k2 = (64, 3, 3) if br < 3 else (128, 5, 7)
        h = stn(c, T, 64, 16, True, True, True) if br < 3 else T
        W, b, pw = c.take(Co, 64, k1, k1), c.take(Co), c.take(1)
        h = F.prelu(F.conv2d(h, W, b, padding=(k1 - 1) // 2), pw)
        h = F.max_pool2d(h, 2) * mk(B * Co).view(B, Co, 1, 1)
        W, b, pw = c.take(Co, Co, k2, k2), c.take(Co), c.take(1)
        outs.append(F.prelu(F.conv2d(h, W, b, padding=(k2 - 1) // 2), pw))
    h = torch.cat(outs, 1) * mk(B * 320).view(B, 320, 1, 1)
    W, b, pw = c.take(256, 20480), c.take(256), c.take(1)
    h = F.prelu(F.linear(h.reshape(B, 20480), W, b), pw) * mk(B * 256).view(B, 256)
    W, b = c.take(1, 256), c.take(1)
    pre = F.linear(h, W, b).view(B)
    assert c.o == flat.numel(), (c.o, flat.numel())
    return torch.sigmoid(pre), pre


def V_forward(flat, running, x, C=3):
    """models.lua:765-804 in evaluate() mode, written with torch.nn.functional (independent of the oracle's operators)."""
    c, r = Cursor(flat), Cursor(running)
    B = x.shape[0]

    def bn(h, ch):
        g, b, m, v = c.take(ch), c.take(ch), r.take(ch), r.take(ch)
        return F.batch_norm(h, m, v, g, b, training=False, eps=1e-5)

    def conv(h, ci, co):
        W, b = c.take(co, ci, 3, 3), c.take(co)
        return F.conv2d(h, W, b, padding=1)

    h = F.max_pool2d(F.leaky_relu(conv(x, C, 128), 0.333), 2)
    h = F.max_pool2d(F.leaky_relu(bn(conv(h, 128, 128), 128), 0.333), 2)
    h = F.leaky_relu(conv(h, 128, 256), 0.333)
    h = F.max_pool2d(F.leaky_relu(bn(conv(h, 256, 256), 256), 0.333), 2) * 0.5
    h = h.reshape(B, 4096)
    for _ in range(2):
        W, b = c.take(1024, h.shape[1]), c.take(1024)
        h = F.leaky_relu(bn(F.linear(h, W, b), 1024), 0.333)
    W, b = c.take(2, 1024), c.take(2)
    h = F.linear(h, W, b)
    assert c.o == flat.numel() and r.o == running.numel()
    return torch.softmax(h, 1)


def bce(p, t, eps=1e-12):
    """nn.BCECriterion, SURVEY.md A.7 (eps inside the log, mean over elements)."""
    return -(t * torch.log(p + eps) + (1 - t) * torch.log(1 - p + eps)).mean()
