"""One build + full statistics of 1e9 points on ONE B200 (180 GB): the n >= 2^30 paths (64-bit look-back words) at full size.
    python tools/big_run.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vapc_b200 import engine as E  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    scale = max(1, n // 100_000_000)  # the density of the bench line: 50 m x 50 m x 4 m per 1e8 points
    cols = []
    for hi in (500_000 * scale, 500_000, 40_000):
        q = torch.randint(0, hi, (n,), device="cuda", generator=g, dtype=torch.int32)
        cols.append(q.to(torch.float64) * 1e-4)
        del q
    torch.cuda.synchronize()
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        c = E.VoxelCloud.build(*cols, 0.1, (0.0, 0.0, 0.0), keep_columns=False)
        st = c.stats(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        total = int(c.count.to(torch.int64).sum().item())
        p2v_ok = bool((c.count.to(torch.int64)[(c.point2voxel[:: 1000].to(torch.int64) & 0xFFFFFFFF)] > 0).all())
        print(f"n={n:.1e}: {ms:.1f} ms = {n / ms / 1e6:.2f} Gpts/s, {c.nvoxels} voxels, {c.grid.total_bits} key bits, sum(count) == n: {total == n}, "
              f"point2voxel rows valid: {p2v_ok}, finite centroids: {bool(torch.isfinite(st.cog).all())}, "
              f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        del c, st


if __name__ == "__main__":
    main()
