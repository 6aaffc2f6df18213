"""Development tool (GPU box): cProfile of the whole facade pipeline (DataFrame in -> compute -> reduce_to_voxels -> DataFrame out).
    python tools/facade_profile.py [n] [return_at]"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.facade_bench import FULL, make_frame, run  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "closest_to_center_of_gravity"
df = make_frame(n)
for _ in range(2):
    print(mode, run(df, FULL, mode))
pr = cProfile.Profile()
pr.enable()
print(run(df, FULL, mode))
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
