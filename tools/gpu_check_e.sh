#!/usr/bin/env bash
# development helper: what the driver runs at round end on a 1-GPU box: the GPU suite, smoke(), the reference arm, the bench line
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo ref rc=$?; cut -c1-400 gpurun_out/bench_r2_ref.json
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo bench rc=$?; tail -2 gpurun_out/bench_r2_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_final.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"])
print("extra", {k: (v["ms_per_step"], v["value"]) for k, v in d["extra"].items()})
print("e2e", {k: v for k, v in d["e2e"].items() if k not in ("api", "platform_what", "facade", "las_int32")})
print("facade", {k: v for k, v in d["e2e"]["facade"].items() if k not in ("api", "unit")})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "launches", d["gpu_launches"], "clocks", d["clocks"])
PY
python tools/facade_bench.py 1e7 2>&1 | tail -8
