#!/bin/bash
# Round 2, first GPU session (1 GPU): whole GPU suite, compute-sanitizer on the tiny kernel cases,
# ncu --set full of the float64 CSR half-products at cfg4 (baseline before the rewrite), bench with parity.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu suite"; timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?"; tail -8 gpurun_out/r2_gpu_tests.log
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout -k 10 420 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest tests/test_gpu_x2.py tests/test_gpu_kernels.py -x -q \
    -k "10-10 or (test_csr_half_plain and 37) or test_slice_rows_key or (test_x2_counts_exact and 129)" \
    > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "rc=$?"; tail -6 gpurun_out/r2_sanitize_$tool.log
done
echo "== ncu full, CSR path"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:csr_half -s 2 -c 2 -o gpurun_out/r2_prof_csr_base -f python bench.py --mode csr --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_ncu_csr_base.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_ncu_csr_base.log
echo "== bench i8 with parity"; timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
echo "rc=$?"; cut -c1-3000 gpurun_out/r2_bench_n1_a.json; tail -3 gpurun_out/r2_bench_n1_a.err
ls -la gpurun_out | tail -12
