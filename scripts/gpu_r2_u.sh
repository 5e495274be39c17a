#!/bin/bash
# second half as ACCUM + FINISH: kernel tests, logical-shard test, then one rank's cfg5 / cfg4 shapes both ways
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 500 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_local_cluster.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -5 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
export SRK_REAL_CFG5=1
export SRK_SWEEP="SRK_FINAL_VIA_ACCUM=0;SRK_FINAL_VIA_ACCUM=1"
timeout -k 10 700 python scripts/csr_shape_bench.py cfg5_s1_final cfg5_s2_final cfg4_n8_final 2>&1 | tee gpurun_out/r2_csr_shapes_finish.jsonl | cut -c1-330
