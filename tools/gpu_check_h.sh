#!/usr/bin/env bash
# development helper (GPU box): reduce_moments with the long / middle segment tiers at several thresholds (build/variants/lib_*.so)
for V in I16 I12; do
  if [ $V = A ]; then unset VAPC_B200_LIB; else export VAPC_B200_LIB=$PWD/build/variants/lib_$V.so; fi
  for W in configs1 configs2; do
    timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/bench_h_${V}_$W.json 2> gpurun_out/bench_h_${V}_$W.err || tail -3 gpurun_out/bench_h_${V}_$W.err
    python - $V $W <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_h_{sys.argv[1]}_{sys.argv[2]}.json"))
k = d["kernels"]
print(sys.argv[1], sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "instrumented", round(d["ms_per_step_instrumented"], 4), "reduce_moments", k["reduce_moments_kernel"]["ms_per_step"], "onesweep", k["onesweep_kernel"]["ms_per_step"], "scatter", k["tile_scatter_kernel"]["ms_per_step"],
      "steps", min(d["step_ms"]), max(d["step_ms"]))
PY
  done
done
for V in I16; do
  export VAPC_B200_LIB=$PWD/build/variants/lib_$V.so
  timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -2
done
