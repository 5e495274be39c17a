#!/bin/bash
# row-sharded first half as ACCUM + FINISH_FIRST: logical-shard tests, then one rank's first-half shapes both ways
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_local_cluster.py -x -q > gpurun_out/r2_local_cluster_tests.log 2>&1
rc=$?; tail -3 gpurun_out/r2_local_cluster_tests.log; echo "tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
export SRK_REAL_CFG5=1
export SRK_SWEEP="SRK_FIRST_VIA_ACCUM=0;SRK_FIRST_VIA_ACCUM=1"
timeout -k 10 300 python scripts/csr_shape_bench.py cfg4_n8_first cfg5_s1_first 2>&1 | tee gpurun_out/r2_csr_shapes_first_via_accum.jsonl | cut -c1-150
