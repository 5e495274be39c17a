#!/bin/bash
# split neighbour lists (SRK_CSR_ACCUM): kernel tests, then the cfg5 shapes of one rank with and without the split
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== csr kernel tests"; timeout -k 5 400 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_kernels.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -5 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
export SRK_SWEEP="SRK_SPLIT_MIN=0;SRK_SPLIT_MIN=2048,SRK_SPLIT_PIECE=1024,SRK_SPLIT_RANGE_MB=32;SRK_SPLIT_MIN=2048,SRK_SPLIT_PIECE=1024,SRK_SPLIT_RANGE_MB=1000;SRK_SPLIT_MIN=256,SRK_SPLIT_PIECE=512,SRK_SPLIT_RANGE_MB=32;SRK_SPLIT_MIN=256,SRK_SPLIT_PIECE=512,SRK_SPLIT_RANGE_MB=16;SRK_SPLIT_MIN=64,SRK_SPLIT_PIECE=256,SRK_SPLIT_RANGE_MB=24"
timeout -k 10 700 python scripts/csr_shape_bench.py cfg5_s2_first cfg5_s2_final cfg5_s1_final 2>&1 | tee -a gpurun_out/r2_csr_split_shapes.jsonl | cut -c1-420
