#!/usr/bin/env bash
# development helper (GPU box): kernel parity tests + a short bench with the per-kernel table (round-2 tile bucket layout)
timeout 1200 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -8
TAG=${1:-f}
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r2_$TAG.json 2> gpurun_out/bench_r2_$TAG.err; echo bench rc=$?; tail -3 gpurun_out/bench_r2_$TAG.err
python - $TAG <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_r2_{sys.argv[1]}.json"))
print(d["ms_per_step"], d["roofline"])
for k, v in d["kernels"].items():
    print(k, v)
for k, v in d["abi_calls"].items():
    print(k, v["ms_per_step"])
print({k: (v["ms_per_step"]) for k, v in d["extra"].items() if isinstance(v, dict) and "ms_per_step" in v})
PY
