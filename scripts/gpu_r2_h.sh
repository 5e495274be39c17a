#!/bin/bash
# Round 2: whole GPU suite (incl. on-device CSR build, auto mode, top-k vs oracle), default bench
# (dense path + CSR path + e2e + cpu baseline), ncu of the csr16 kernels for profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu suite"; timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?"; tail -12 gpurun_out/r2_gpu_tests.log
echo "== bench default"; timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "rc=$?"; cut -c1-6000 gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
echo "== ncu full csr16"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"csr_gather|quantize|row_unit" -s 4 -c 4 -o gpurun_out/r2_prof_csr16 -f python bench.py --mode csr16 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_ncu_csr16.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_ncu_csr16.log
