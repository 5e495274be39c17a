#!/bin/bash
# Round 2, multi-GPU parity session (gpurun --gpus 2|4|8): both row-sharded solvers against the oracle.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
echo "== row-sharded parity on $n GPUs"
timeout -k 10 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_zz_dist_csr.py -x -q > gpurun_out/r2_dist_tests_n$n.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/r2_dist_tests_n$n.log
