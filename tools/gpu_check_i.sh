#!/usr/bin/env bash
# development helper (GPU box): reduce_moments with 1 / 2 / 4 lanes per voxel on the bench cloud (10 points per voxel) and at 1.0 m voxels
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from tools import synth
from vapc_b200 import engine as E, _lib as L
dev = torch.device("cuda", 0)
x, y, z = synth.uniform_cloud(100_000_000, seed=2, device=dev, extent=(50.0, 50.0, 4.0))
for vs in (0.1, 0.2, 0.05):
    for lanes in (1, 2, 4, 0):
        L.call("vapc_reduce_debug_lanes", lanes)
        for _ in range(3):
            c = E.VoxelCloud.build(x, y, z, vs, (0.0, 0.0, 0.0), keep_columns=False); c.point2voxel; c = None
        torch.cuda.synchronize()
        L.call("vapc_profile_enable", 1)
        for _ in range(5):
            c = E.VoxelCloud.build(x, y, z, vs, (0.0, 0.0, 0.0), keep_columns=False); c.point2voxel; v = c.nvoxels; c = None
        torch.cuda.synchronize()
        L.call("vapc_profile_enable", 0)
        kt = L.profile_collect()
        ms = sum(m for k, (cnt, m) in kt.items() if "reduce_moments" in k) / 5
        print(f"voxel {vs} m ({100e6 / v:.1f} points per voxel): lanes {lanes}: reduce_moments {ms:.4f} ms", flush=True)
L.call("vapc_reduce_debug_lanes", 0)
PY
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -2
