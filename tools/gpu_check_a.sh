#!/usr/bin/env bash
# development helper (runs on the GPU box through gpurun): cell-sort tests + a short bench with the per-kernel table
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "cell_sort or voxel_table or reproducible or properties" 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo bench rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_a.json"))
print(d["ms_per_step"], d["roofline"])
for k, v in d["kernels"].items():
    print(k, v)
print(d["gaps_between_calls_ms"])
PY
