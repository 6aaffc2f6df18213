#!/usr/bin/env bash
# development helper (GPU box): point->voxel lookups on a side stream: step time against the lookups' CTAs per SM
for C in off 2 3 4 6 8; do
  if [ $C = off ]; then export VAPC_P2V_SIDE_STREAM=0; else export VAPC_P2V_SIDE_STREAM=1 VAPC_P2V_SIDE_CTAS=$C; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r2_g_$C.json 2> gpurun_out/bench_r2_g_$C.err || tail -3 gpurun_out/bench_r2_g_$C.err
  python - $C <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_r2_g_{sys.argv[1]}.json"))
k = d["kernels"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "p2v_lookup", k["p2v_lookup_kernel"]["ms_per_step"], "stats", k["voxel_stats_kernel"]["ms_per_step"],
      "configs2", round(d["extra"]["configs2_one_gpu"]["ms_per_step"], 3), "bounds_computed", round(d["extra"]["bounds_computed"]["ms_per_step"], 3))
PY
done
unset VAPC_P2V_SIDE_STREAM VAPC_P2V_SIDE_CTAS
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_facade.py -x -q -m gpu 2>&1 | tail -4
