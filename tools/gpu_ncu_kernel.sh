#!/usr/bin/env bash
# development helper: one `ncu --set full` capture of a single kernel of the bench step.   usage: gpu_ncu_kernel.sh <kernel regex> <tag>
K=${1:-cell_sort_kernel}; TAG=${2:-r2_a}
VAPC_CELL_SORT=${VAPC_CELL_SORT:-0} python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_${TAG}_plain.err || { echo "bench failed"; tail -5 gpurun_out/ncu_${TAG}_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:${K} -s 3 -c 1 -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
