#!/bin/bash
# compute-sanitizer memcheck over the ACCUM / FINISH cases after the cp.async-staged streaming passes
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 compute-sanitizer --tool memcheck --error-exitcode 9 \
  python -m pytest tests/test_gpu_csr_gather.py -x -q -k "finish or accum or split or via_accum" \
  > gpurun_out/r2_sanitize_csr_stream_memcheck.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/r2_sanitize_csr_stream_memcheck.log
