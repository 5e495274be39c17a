#!/bin/bash
# compute-sanitizer passes over the small kernel tests (memcheck, racecheck, synccheck).  Slow under
# the tool: only the cases with tiny shapes are selected.  NOT yet run in round 1 (GPU budget spent).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout -k 10 600 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest tests/test_gpu_x2.py tests/test_gpu_kernels.py -x -q \
    -k "10-10 or (test_csr_half_plain and 37) or test_slice_rows_key or test_x2_counts_exact" \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; tail -5 gpurun_out/sanitize_$tool.log
done
