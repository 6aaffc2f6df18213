#!/usr/bin/env bash
# development helper: N-GPU bench line (+ the N-rank parity tests when asked).   usage: gpu_check_n.sh <N> [tests]
N=${1:-2}
if [ "${2:-}" = "tests" ]; then timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | grep -v "^E   *$" | grep -B2 -A12 "Traceback\|Error\|assert" | head -80; fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
  > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo bench rc=$?; tail -15 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_r2_n$N.json"))
except Exception as e:
    print("no json", e); raise SystemExit
print("ms/step", d["ms_per_step"], "value", d["value"], "voxels", d["config"].get("voxels_all_gpus"), "key_bits", d["config"]["key_bits"])
print("check", d["check"])
print("extra", json.dumps(d["extra"], indent=1))
print("e2e", {k: v for k, v in (d["e2e"] or {}).items() if k not in ("api", "platform_what")})
for k, v in d["kernels"].items():
    print(k, v["launches_per_step"], v["ms_per_step"], v["GBps"])
print(d["abi_calls"].keys())
PY
