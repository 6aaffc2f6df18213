"""Timing of the neighbour queries on the GPU (CUDA events): nearest neighbours between two jittered centroid sets and
connected components of occupied voxels.    python tools/nbench.py [n]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from vapc_b200 import engine as E  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    g = torch.Generator(device="cuda").manual_seed(1)
    ext = torch.tensor([500.0, 500.0, 20.0], device="cuda", dtype=torch.float64)
    a = torch.rand((3, n), generator=g, device="cuda", dtype=torch.float64) * ext[:, None]
    b = a + (torch.rand((3, n), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 0.1
    h = float((500 * 500 * 20 / n) ** (1 / 3))
    ms, cells = timed(lambda: E._cell_cloud(b[0], b[1], b[2], h))
    print(f"cell list over {n} points (cell {h:.3f} m, {cells.nvoxels} cells): {ms:.2f} ms")
    ms, (idx, dist) = timed(lambda: E.nearest_neighbors(a[0], a[1], a[2], cells))
    print(f"vapc_cell_nn {n} queries: {ms:.2f} ms = {n / ms / 1e3:.1f} M queries/s; mean distance {float(dist.mean()):.4f}")
    t = time.time()
    ms, (i1, i2, d) = timed(lambda: E.mutual_nearest_neighbors(list(a), list(b), cell_sizes=(h, h)), reps=1)
    print(f"mutual nearest neighbours {n} x {n}: {ms:.1f} ms, {i1.numel()} pairs")
    # occupied voxels of a sparse grid, 6-neighbourhood
    side = int(round((n / 0.25) ** (1 / 3)))
    rows = torch.randint(0, side, (3, n), generator=g, device="cuda").to(torch.float64)
    for dist_v in (1.0, 1.8):
        ms, (label, size, opens) = timed(lambda: E.cluster_components(rows[0], rows[1], rows[2], dist_v, integer_rows=True), reps=1)
        print(f"cluster_components {n} rows in {side}^3 voxels, distance {dist_v}: {ms:.1f} ms, {int(label.max())} labels opened {opens}, largest {int(size.max())}")


if __name__ == "__main__":
    main()
