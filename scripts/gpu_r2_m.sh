#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_local_cluster.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -8 gpurun_out/r2_csr_tests.log; echo "tests rc=$rc"
