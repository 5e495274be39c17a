#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export SRK_REAL_CFG5=1
timeout -k 10 700 python scripts/csr_shape_bench.py cfg5_s1_final cfg5_s2_final cfg5_s2_first cfg5_s1_first cfg4_n8_final 2>&1 | tee gpurun_out/r2_csr_shapes_batch8.jsonl | cut -c1-330
