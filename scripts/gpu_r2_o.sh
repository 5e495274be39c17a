#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== csr kernel tests"; timeout -k 5 400 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_kernels.py tests/test_gpu_local_cluster.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -5 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
for tc in 512 256; do
  echo "== shapes, SRK_CSR_TC=$tc"
  SRK_CSR_TC=$tc timeout -k 10 600 python scripts/csr_shape_bench.py cfg4_n8_final cfg4_n8_first cfg5_s1_final cfg5_s2_first cfg5_s2_final cfg5_s1_first 2>&1 | tee -a gpurun_out/r2_csr_shapes.jsonl | cut -c1-400
done
