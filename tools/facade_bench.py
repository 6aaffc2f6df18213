"""The drop-in `Vapc` facade timed end to end on the GPU box: pandas DataFrame in -> (reduced) DataFrame out, host time included.

    python tools/facade_bench.py [n]            table of the usual pipelines at n points (default 1e7)
    facade_e2e(sizes)                           bench.py's `e2e.facade` entry: full attribute set -> reduce_to_voxels

The frame is what the reference's DataHandler hands over: X, Y, Z + two attribute columns, all float64 (datahandler.py:83-88).
"""
import os
import sys
import time

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vapc_b200 as vapc  # noqa: E402

FULL = ["point_count", "point_density", "center_of_gravity", "std_of_cog", "covariance_matrix", "eigenvalues", "geometric_features"]


def make_frame(n, seed=0):
    rng = np.random.default_rng(seed)
    return pd.DataFrame({"X": rng.integers(0, 500000, n) * 1e-4, "Y": rng.integers(0, 500000, n) * 1e-4, "Z": rng.integers(0, 40000, n) * 1e-4,
                         "intensity": rng.integers(0, 1000, n).astype(float), "classification": rng.integers(1, 9, n).astype(float)})


def run(df, compute, reduce_to=None, stats=None, voxel_size=0.1, read_frame=True):
    """One pipeline on a fresh Vapc over a copy of df; returns (seconds, shape of the resulting frame)."""
    v = vapc.Vapc(voxel_size, [0, 0, 0])
    v.df = df.copy()
    v.original_attributes = list(v.df.columns)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if compute:
        v.compute = compute
        v.compute_requested_attributes()
    if stats:
        v.attributes = stats
        v.compute_requested_statistics_per_attributes()
    if reduce_to:
        v.return_at = reduce_to
        v.reduce_to_voxels()
    out = v.df if read_frame else None  # reading the frame is part of the job: pending columns are materialised here
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt, (out.shape if out is not None else None)


def facade_e2e(sizes=(10_000_000,), reps=2):
    """bench.py's e2e.facade: DataFrame in -> full attribute set -> reduce_to_voxels -> DataFrame out, per size and reduction point."""
    out = {"api": "vapc_b200.Vapc (the reference's class): v.df = DataFrame(X, Y, Z, intensity, classification); v.compute = full attribute set; "
                  "v.compute_requested_attributes(); v.reduce_to_voxels(); v.df   (wall clock incl. host -> device of X, Y, Z and the reduced frame "
                  "back, best of %d)" % reps, "unit": "points/s"}
    for n in sizes:
        df = make_frame(n)
        run(df.iloc[: min(n, 200_000)], FULL, "center_of_voxel")  # warm-up (allocators, pinned buffers)
        for mode in ("center_of_voxel", "closest_to_center_of_gravity"):
            best = min(run(df, FULL, mode)[0] for _ in range(reps))
            out[f"{n:.0e}".replace("+0", "") + "_" + mode] = {"seconds": round(best, 4), "value": n / best}
        del df
    return out


if __name__ == "__main__":
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    df = make_frame(n)
    for name, args in (("warmup", (["point_count"],)), ("point_count", (["point_count"],)),
                       ("count+cog+cov", (["point_count", "center_of_gravity", "covariance_matrix"],)),
                       ("full per point", (FULL,)), ("full -> center_of_voxel", (FULL, "center_of_voxel")),
                       ("full -> closest", (FULL, "closest_to_center_of_gravity")),
                       ("stats mode/mean", ([], None, {"classification": ["mode", "mode_count,0.1", "mean"]}))):
        dt, shape = run(df, *args)
        print(f"{name:28s} {dt * 1e3:9.1f} ms  {n / dt / 1e6:8.2f} Mpts/s  df {shape}", flush=True)
    if len(sys.argv) > 2 and sys.argv[2] == "profile":
        import cProfile
        import pstats

        v = vapc.Vapc(0.1, [0, 0, 0])
        v.df = df.copy()
        v.original_attributes = list(v.df.columns)
        v.compute = FULL
        v.compute_requested_attributes()
        v.return_at = sys.argv[3] if len(sys.argv) > 3 else "center_of_voxel"
        pr = cProfile.Profile()
        pr.enable()
        v.reduce_to_voxels()
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
