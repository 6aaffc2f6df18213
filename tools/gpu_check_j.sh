#!/usr/bin/env bash
# development helper (GPU box): tile_scatter_kernel at other tile sizes / CTAs per SM (build/variants/lib_S*.so: items x min blocks)
for V in A S44 S45 S64 S83; do
  if [ $V = A ]; then unset VAPC_B200_LIB; else export VAPC_B200_LIB=$PWD/build/variants/lib_$V.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/bench_j_$V.json 2> gpurun_out/bench_j_$V.err || tail -3 gpurun_out/bench_j_$V.err
  python - $V <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_j_{sys.argv[1]}.json"))
k = d["kernels"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "scatter", k["tile_scatter_kernel"]["ms_per_step"], "digits", k["tile_digits_kernel"]["ms_per_step"], "offsets", k["tile_offsets_kernel"]["ms_per_step"], "reduce", k["reduce_moments_kernel"]["ms_per_step"])
PY
done
