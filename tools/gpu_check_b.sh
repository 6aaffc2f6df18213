#!/usr/bin/env bash
# development helper: the whole GPU suite + the N = 1 bench line + the facade table
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; echo bench rc=$?; tail -3 gpurun_out/bench_r2_b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_b.json"))
print("ms/step", d["ms_per_step"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "extra", d["extra"])
print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("api", "platform_what")})
print("cpu", d["cpu_baseline"])
for k, v in d["kernels"].items():
    print(k, v["ms_per_step"], v["GBps"])
PY
python tools/facade_bench.py 1e7 2>&1 | tail -12
