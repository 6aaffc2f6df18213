"""Synthetic clouds of the BASELINE.json configurations, generated on the device.

    uniform_cloud      configs[1]: quantised uniform points in a box
    clustered_chunk    configs[2]: one rank's contiguous chunk of a terrain + trees cloud whose y-extent grows with the number of
                       ranks (weak scaling at constant density), in scan-line order (rank r holds the strip y in [r W, (r+1) W):
                       spatially coherent chunks, like consecutive flight lines of a LAS file) or shuffled (every rank holds a
                       random sample of the WHOLE cloud: the worst case for the exchange, every rank touches every voxel)

The tree list is a function of the global seed only, so every rank places the same trees; the points are drawn per rank.
"""
import math

import torch

QUANT = 1e-4  # LAS-like coordinate quantisation
STRIP_WIDTH = 125.0  # metres of y per rank at 1.25e8 points: 1000 m x 125 m per GPU = the density of 1e9 points per km^2
EXTENT_X = 1000.0
TREES_PER_M2 = 0.2  # 2e5 trees per km^2


def uniform_cloud(n, seed, device, extent=(50.0, 50.0, 4.0), x_shift=0.0):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cols = []
    for k, ext in enumerate(extent):
        q = torch.randint(0, int(round(ext / QUANT)), (n,), generator=g, device=device, dtype=torch.int64)
        c = q.to(torch.float64) * QUANT
        if k == 0 and x_shift:
            c += x_shift
        cols.append(c)
        del q
    return cols


def _terrain(x, y):
    return 5.0 * torch.sin(x / 37.0) * torch.cos(y / 53.0) + 0.02 * x


def _trees(seed, y0, y1, device, extent_x=EXTENT_X):
    """Tree centres and heights with y in [y0, y1): drawn strip by strip (10 m) from the global seed, so that any rank asking for
    any y-range sees the same trees."""
    s0, s1 = int(math.floor(y0 / 10.0)), int(math.ceil(y1 / 10.0))
    per = max(1, int(TREES_PER_M2 * extent_x * 10.0))
    cx, cy, ch = [], [], []
    for s in range(s0, s1):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed * 1_000_003 + s)
        u = torch.rand(3, per, generator=g, dtype=torch.float64)
        cx.append(u[0] * extent_x)
        cy.append((s + u[1]) * 10.0)
        ch.append(5.0 + 20.0 * u[2])
    cx, cy, ch = (torch.cat(v).to(device) for v in (cx, cy, ch))
    keep = (cy >= y0) & (cy < y1)
    return cx[keep], cy[keep], ch[keep]


def clustered_chunk(n, rank, world, device, seed=3, scan_order=True, strip_width=None, extent_x=EXTENT_X):
    """This rank's n points of the configs[2] cloud over world ranks: 60 % terrain sheet z = f(x, y) +- 5 cm, 40 % trees (thin stems +
    Gaussian crowns, height 5 .. 25 m), x in [0, extent_x = 1000) m, y in [0, world * strip_width) m; quantised to 1e-4 m.  The density
    (1e6 points per 1000 m2) is kept whatever n: the default strip width follows from n and extent_x (a crop for the CPU legs: extent_x =
    strip width = sqrt(n / 1000) m)."""
    w = float(strip_width if strip_width is not None else STRIP_WIDTH * n / 1.25e8 * (EXTENT_X / extent_x))
    if scan_order:
        y0, y1 = rank * w, (rank + 1) * w
    else:
        y0, y1 = 0.0, world * w
    g = torch.Generator(device=device)
    g.manual_seed(seed * 7919 + 31 * rank + (0 if scan_order else 1_000_000))
    nt = int(0.6 * n)
    tx = torch.rand(nt, device=device, generator=g, dtype=torch.float64) * extent_x
    ty = y0 + torch.rand(nt, device=device, generator=g, dtype=torch.float64) * (y1 - y0)
    tz = _terrain(tx, ty) + (torch.rand(nt, device=device, generator=g, dtype=torch.float64) - 0.5) * 0.1
    nb = n - nt
    cx, cy, ch = _trees(seed, y0, y1, device, extent_x)
    if cx.numel() == 0:  # a crop smaller than the tree spacing: one tree in its middle keeps the 60 / 40 mix
        cx = torch.tensor([0.5 * extent_x], dtype=torch.float64, device=device)
        cy = torch.tensor([0.5 * (y0 + y1)], dtype=torch.float64, device=device)
        ch = torch.tensor([15.0], dtype=torch.float64, device=device)
    t = torch.randint(0, cx.numel(), (nb,), device=device, generator=g)
    r = torch.randn(nb, 3, device=device, generator=g, dtype=torch.float64)
    crown = torch.rand(nb, device=device, generator=g) < 0.8
    bx = cx[t] + torch.where(crown, r[:, 0] * 1.5, r[:, 0] * 0.1)
    by = cy[t] + torch.where(crown, r[:, 1] * 1.5, r[:, 1] * 0.1)
    h = torch.rand(nb, device=device, generator=g, dtype=torch.float64)
    ground = _terrain(cx[t], cy[t])
    bz = ground + torch.where(crown, ch[t] * (0.6 + 0.4 * h) + r[:, 2] * 0.8, ch[t] * 0.6 * h)
    del r, t, h, crown, ground
    x, y, z = torch.cat([tx, bx]), torch.cat([ty, by]), torch.cat([tz, bz])
    del tx, ty, tz, bx, by, bz
    x, y, z = (torch.round(c / QUANT) * QUANT for c in (x, y, z))
    if scan_order:  # scan lines: 2 m strips in y, x ascending inside a strip
        key = torch.floor(y / 2.0) * 1e6 + x
        perm = torch.argsort(key)
        del key
    else:
        perm = torch.randperm(n, device=device, generator=g)
    out = tuple(c[perm].contiguous() for c in (x, y, z))
    del perm
    return out
