#!/usr/bin/env bash
# development helper: the scale tests, then the N-rank parity worker + bench at N GPUs
N=${1:-4}
timeout 900 python -m pytest tests/test_gpu_scale.py -x -q -m gpu 2>&1 | tail -4
bash tools/gpu_check_n.sh $N tests
