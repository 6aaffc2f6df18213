#!/usr/bin/env bash
# development helper (GPU box): reduce_moments_kernel with the records of the tile N tiles ahead requested into L2 (build/variants/lib_R*.so)
for V in A R296 R592 R1184; do
  if [ $V = A ]; then unset VAPC_B200_LIB; else export VAPC_B200_LIB=$PWD/build/variants/lib_$V.so; fi
  for W in configs1 configs2; do
    timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/bench_k_${V}_$W.json 2> gpurun_out/bench_k_${V}_$W.err || tail -3 gpurun_out/bench_k_${V}_$W.err
    python - $V $W <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_k_{sys.argv[1]}_{sys.argv[2]}.json"))
k = d["kernels"]
print(sys.argv[1], sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "reduce", k["reduce_moments_kernel"]["ms_per_step"], "scatter", k["tile_scatter_kernel"]["ms_per_step"])
PY
  done
done
