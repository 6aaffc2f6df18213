#!/usr/bin/env bash
# development helper: cluster cell sort + facade tests, then timing of the cell-sort step
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_facade.py -x -q -m gpu 2>&1 | tail -25
VAPC_CELL_SORT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err; echo bench rc=$?; tail -3 gpurun_out/bench_r2_c.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_c.json"))
print("ms/step", d["ms_per_step"], d["config"].get("grouping"), "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "extra", d["extra"])
for k, v in d["kernels"].items():
    print(k, v["ms_per_step"], v["GBps"])
print(d["gaps_between_calls_ms"])
PY
