"""Timing of the BASELINE.json configurations that are not the bench line (configs[2..4]: parity-test cases in tests/test_gpu_scale.py),
one GPU, CUDA events, inputs resident.    python tools/configs_bench.py [n]   (default 1e8 points; the two-epoch case uses 2 x 2n)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.kbench import clustered_cloud  # noqa: E402
from vapc_b200 import engine as E  # noqa: E402


def timed(fn, reps=3, warm=2):
    """Best of `reps` single-call CUDA-event timings after `warm` calls (the first calls of a new size grow torch's caching
    allocator with cudaMalloc)."""
    out = None
    for _ in range(warm):
        out = None
        out = fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        out = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


FULL = dict(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)


def full_step(x, y, z, vs):
    c = E.VoxelCloud.build(x, y, z, vs, (0.0, 0.0, 0.0))
    c.stats(**FULL)
    return c


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    g = torch.Generator(device="cuda").manual_seed(2)
    # configs[2] per GPU: clustered terrain + trees, scan-line order and shuffled
    for order in (True, False):
        x, y, z = clustered_cloud(n, scan_order=order)
        for vs in (0.25, 0.1):
            ms, c = timed(lambda: full_step(x, y, z, vs))
            print(f"configs[2] clustered {'scan-line' if order else 'shuffled '} n={n:.0e} vs={vs}: {ms:8.2f} ms = {n / ms / 1e6:6.2f} Gpts/s  "
                  f"({c.nvoxels} voxels, {c.grid.total_bits} key bits)", flush=True)
            del c
        del x, y, z
    # configs[4]: voxel-size sweep on the uniform cloud with an integer class attribute: full set, mode + mode_count, closest point
    x = torch.randint(0, 500_000, (n,), device="cuda", generator=g).to(torch.float64) * 1e-4
    y = torch.randint(0, 500_000, (n,), device="cuda", generator=g).to(torch.float64) * 1e-4
    z = torch.randint(0, 40_000, (n,), device="cuda", generator=g).to(torch.float64) * 1e-4
    cls = torch.randint(1, 9, (n,), device="cuda", generator=g).to(torch.float64)
    for vs in (0.05, 0.2, 1.0):
        ms, c = timed(lambda: full_step(x, y, z, vs))
        ms_mode, _ = timed(lambda: c.mode_count(cls, [0.1], want_mode=True))
        ms_close, _ = timed(lambda: c.closest_to_cog())
        print(f"configs[4] uniform n={n:.0e} vs={vs}: full set {ms:7.2f} ms = {n / ms / 1e6:6.2f} Gpts/s ({c.nvoxels} voxels, {c.grid.total_bits} key bits); "
              f"mode + mode_count,0.1 {ms_mode:6.2f} ms; closest-to-centroid {ms_close:6.2f} ms", flush=True)
        del c
    del x, y, z, cls
    # configs[3]: two epochs of 2n points, voxel join + centroid distance + Mahalanobis test
    m = 2 * n
    e1 = clustered_cloud(m, seed=3, extent=700.0)
    noise = [torch.randn(m, device="cuda", generator=g, dtype=torch.float64) * 0.01 for _ in range(3)]
    e2 = [torch.round((c + d) * 1e4) * 1e-4 for c, d in zip(e1, noise)]
    del noise
    moved = (e2[0] > 300.0) & (e2[0] < 335.0)
    e2[0] = torch.where(moved, e2[0] + 0.3, e2[0])
    for vs in (0.5,):
        def change():
            tabs = []
            for e in (e1, e2):
                c = E.VoxelCloud.build(*e, vs, (0.0, 0.0, 0.0), want_point2voxel=False, keep_columns=False)
                tabs.append((c, c.stats(cog=True, cov=True)))
            i1, i2, only1, only2 = E.join_voxel_tables(tabs[0][0].voxel_coords(), tabs[1][0].voxel_coords())
            res = E.mahalanobis(tabs[0][1].cog, tabs[0][1].cov, tabs[1][1].cog, tabs[1][1].cov, i1, i2, alpha=0.005)
            return i1.numel(), only1.numel(), only2.numel(), int((res["change_type"] > 0).sum())
        ms, out = timed(change, reps=2)
        print(f"configs[3] two epochs of {m:.0e} points vs={vs}: {ms:8.2f} ms = {2 * m / ms / 1e6:6.2f} Gpts/s  (joined {out[0]}, disappeared {out[1]}, "
              f"appeared {out[2]}, significant {out[3]})", flush=True)


if __name__ == "__main__":
    main()
