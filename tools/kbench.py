"""Kernel micro-benchmarks on the GPU box (development tool, not part of the product or the tests).

    python tools/kbench.py sort [n] [bits]      stable pair sort vs torch.sort, CUDA-event timing
    python tools/kbench.py build [n]            the VoxelCloud.build + stats pipeline, per C-ABI call
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vapc_b200 import _lib as L  # noqa: E402
from vapc_b200 import engine as E  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def bench_sort(n=100_000_000, bits=24):
    g = torch.Generator(device="cuda").manual_seed(1)
    if bits <= 32:
        keys = torch.randint(0, 1 << min(bits, 31), (n,), device="cuda", dtype=torch.int64, generator=g).to(torch.int32)
    else:
        keys = torch.randint(0, 1 << min(bits, 62), (n,), device="cuda", dtype=torch.int64, generator=g)
    ks, idx = E.sort_pairs(keys.clone(), bits)
    torch.cuda.synchronize()
    ref_k, ref_i = torch.sort(keys, stable=True)
    ok = bool(torch.equal(ks, ref_k)) and bool(torch.equal(idx.to(torch.int64) & 0xFFFFFFFF, ref_i))
    del ref_k, ref_i, ks, idx
    src = keys.clone()

    def run():
        src.copy_(keys)
        E.sort_pairs(src, bits)

    def copy_only():
        src.copy_(keys)

    t_copy = timed(copy_only)[0]
    best, med = timed(run)
    kb = keys.element_size()
    passes = (bits + 7) // 8
    alg = n * (kb + passes * 2 * (kb + 4) - 4)
    print(f"sort n={n} bits={bits} key_bytes={kb} passes={passes}: correct={ok} best {best - t_copy:.3f} ms (median {med - t_copy:.3f}) "
          f"= {(best - t_copy) / passes:.3f} ms/pass; algorithmic {alg / 1e9:.2f} GB -> {alg / (best - t_copy) / 1e6:.0f} GB/s")
    return ok


def clustered_cloud(n, seed=3, extent=500.0, scan_order=True):
    """C3-like cloud: 60 % of the points on a terrain sheet z = f(x, y) + noise, 40 % in tree-like
    blobs (stems + crowns); quantised to 1e-4 m; emitted in scan-line order (sorted by a coarse y strip, then x) or shuffled."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    nt = int(0.6 * n)
    tx = torch.rand(nt, device="cuda", generator=g, dtype=torch.float64) * extent
    ty = torch.rand(nt, device="cuda", generator=g, dtype=torch.float64) * extent
    tz = 5.0 * torch.sin(tx / 37.0) * torch.cos(ty / 53.0) + 0.02 * tx + (torch.rand(nt, device="cuda", generator=g, dtype=torch.float64) - 0.5) * 0.1
    nb = n - nt
    ntrees = max(1, int(2e5 * (extent / 1000.0) ** 2))
    cx = torch.rand(ntrees, device="cuda", generator=g, dtype=torch.float64) * extent
    cy = torch.rand(ntrees, device="cuda", generator=g, dtype=torch.float64) * extent
    ch = 5.0 + 20.0 * torch.rand(ntrees, device="cuda", generator=g, dtype=torch.float64)
    t = torch.randint(0, ntrees, (nb,), device="cuda", generator=g)
    r = torch.randn(nb, 3, device="cuda", generator=g, dtype=torch.float64)
    crown = torch.rand(nb, device="cuda", generator=g) < 0.8
    bx = cx[t] + torch.where(crown, r[:, 0] * 1.5, r[:, 0] * 0.1)
    by = cy[t] + torch.where(crown, r[:, 1] * 1.5, r[:, 1] * 0.1)
    h = torch.rand(nb, device="cuda", generator=g, dtype=torch.float64)
    ground = 5.0 * torch.sin(cx[t] / 37.0) * torch.cos(cy[t] / 53.0) + 0.02 * cx[t]
    bz = ground + torch.where(crown, ch[t] * (0.6 + 0.4 * h) + r[:, 2] * 0.8, ch[t] * 0.6 * h)
    x, y, z = torch.cat([tx, bx]), torch.cat([ty, by]), torch.cat([tz, bz])
    del tx, ty, tz, bx, by, bz, r, t, h, crown, ground
    x, y, z = (torch.round(c * 1e4) * 1e-4 for c in (x, y, z))
    if scan_order:
        key = torch.floor(y / 2.0) * 1e6 + x
        perm = torch.argsort(key)
    else:
        perm = torch.randperm(n, device="cuda", generator=g)
    return x[perm].contiguous(), y[perm].contiguous(), z[perm].contiguous()


def bench_build(n=100_000_000, bucketed=1, workload=0, vs_mm=100):
    """workload 0: uniform box of bench.py; 1: clustered, scan-line order; 2: clustered, shuffled.  vs_mm: voxel size in mm."""
    g = torch.Generator(device="cuda").manual_seed(2)
    if workload == 0:
        x = torch.randint(0, 500_000, (n,), device="cuda", generator=g).to(torch.float64) * 1e-4
        y = torch.randint(0, 500_000, (n,), device="cuda", generator=g).to(torch.float64) * 1e-4
        z = torch.randint(0, 40_000, (n,), device="cuda", generator=g).to(torch.float64) * 1e-4
    else:
        x, y, z = clustered_cloud(n, scan_order=workload == 1)
    vs = vs_mm / 1000.0
    rec = []

    def obs(name, phase):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        rec.append((name, phase, ev))

    def step():
        c = E.VoxelCloud.build(x, y, z, vs, (0.0, 0.0, 0.0), bucketed=bool(bucketed))
        c.stats(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)
        return c

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    L.observer = obs
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        c = step()
    b.record()
    torch.cuda.synchronize()
    L.observer = None
    per, pend = {}, {}
    for name, phase, ev in rec:
        if phase == "before":
            pend[name] = ev
        else:
            per.setdefault(name, []).append(pend.pop(name).elapsed_time(ev))
    print(f"build+stats workload={workload} vs={vs} key_bits={c.grid.total_bits} bucketed={bucketed} n={n} voxels={c.nvoxels}: {a.elapsed_time(b) / reps:.3f} ms/step = {n * reps / a.elapsed_time(b) / 1e6:.2f} Gpts/s")
    for k, v in per.items():
        print(f"   {k:28s} {sum(v) / reps:.3f} ms/step ({len(v) // reps} calls)")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "sort"
    args = [int(float(a)) for a in sys.argv[2:]]
    t0 = time.time()
    if what == "sort":
        ok = bench_sort(*args)
        sys.exit(0 if ok else 1)
    elif what == "build":
        bench_build(*args)
    print(f"[{time.time() - t0:.1f} s]")
