import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, pandas as pd, torch
import vapc_b200 as vapc
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(0)
df = pd.DataFrame({"X": rng.integers(0, 500000, n) * 1e-4, "Y": rng.integers(0, 500000, n) * 1e-4, "Z": rng.integers(0, 40000, n) * 1e-4,
                   "intensity": rng.integers(0, 1000, n).astype(float), "classification": rng.integers(1, 9, n).astype(float)})
def run(compute, reduce_to=None, stats=None):
    v = vapc.Vapc(0.1, [0, 0, 0])
    v.df = df.copy()
    v.original_attributes = list(v.df.columns)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if compute:
        v.compute = compute
        v.compute_requested_attributes()
    if stats:
        v.attributes = stats
        v.compute_requested_statistics_per_attributes()
    if reduce_to:
        v.return_at = reduce_to
        v.reduce_to_voxels()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return dt, v.df.shape
full = ["point_count", "point_density", "center_of_gravity", "std_of_cog", "covariance_matrix", "eigenvalues", "geometric_features"]
for name, args in (("warmup", (["point_count"],)), ("point_count", (["point_count"],)), ("count+cog+cov", (["point_count", "center_of_gravity", "covariance_matrix"],)),
                   ("full per point", (full,)), ("full -> center_of_voxel", (full, "center_of_voxel")), ("full -> closest", (full, "closest_to_center_of_gravity")),
                   ("stats mode/mean", ([], None, {"classification": ["mode", "mode_count,0.1", "mean"]}))):
    dt, shape = run(*args)
    print(f"{name:28s} {dt*1e3:9.1f} ms  {n/dt/1e6:8.2f} Mpts/s  df {shape}", flush=True)

if len(sys.argv) > 2 and sys.argv[2] == "profile":
    import cProfile, pstats
    v = vapc.Vapc(0.1, [0, 0, 0])
    v.df = df.copy()
    v.original_attributes = list(v.df.columns)
    v.compute = full
    pr = cProfile.Profile()
    pr.enable()
    v.compute_requested_attributes()
    v.return_at = sys.argv[3] if len(sys.argv) > 3 else "center_of_voxel"
    pr.disable()
    pr = cProfile.Profile()  # profile the reduction alone
    pr.enable()
    v.reduce_to_voxels()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
