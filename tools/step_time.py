"""Development tool (GPU box): the bench step (configs[1]) timed without the per-kernel profiling events.
    python tools/step_time.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402
from vapc_b200 import engine as E  # noqa: E402

n = 100_000_000
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
x, y, z = synth.uniform_cloud(n, seed=2, device=dev, extent=(50.0, 50.0, 4.0))
bounds = [0.0, 0.0, 0.0, 50.0 - 1e-4, 50.0 - 1e-4, 4.0 - 1e-4]
FULL = dict(density=True, cog=True, std=True, cov=True, eig=True, eig_n=True, feat=True)


def step():
    c = E.VoxelCloud.build(x, y, z, 0.1, (0.0, 0.0, 0.0), keep_columns=False, bounds=bounds)
    st = c.stats(**FULL)
    c.point2voxel
    return c, st


for _ in range(3):
    out = None
    out = step()
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    out = None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = step()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
out = None
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    out = None
    out = step()
b.record()
torch.cuda.synchronize()
ts.sort()
print(f"side={E.P2V_SIDE_STREAM} ctas={E.P2V_SIDE_CTAS} prio={E.P2V_SIDE_PRIORITY}: single steps min {ts[0]:.4f} median {ts[len(ts) // 2]:.4f} max {ts[-1]:.4f}; "
      f"{reps} back to back {a.elapsed_time(b) / reps:.4f} ms/step, V={out[0].nvoxels}")
