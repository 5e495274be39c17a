#!/bin/bash
# Round 2: TMA gather-rate microbenchmark + parity suite with the csr16 mode.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tma gather rate"; timeout -k 5 300 scripts/micro/tma_gather_rate > gpurun_out/r2_micro_tma_gather_rate.txt 2>&1
echo "rc=$?"; cat gpurun_out/r2_micro_tma_gather_rate.txt
echo "== parity suite"; timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q > gpurun_out/r2_parity_tests.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/r2_parity_tests.log
