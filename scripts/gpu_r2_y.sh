#!/bin/bash
# streaming passes with the one-wave-ahead L2 prefetch (SRK_CSR_FLAGS=8 switches it off): tests, csr16 bench, shapes
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_csr_gather.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -3 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
for fl in 8 0; do
echo "== bench csr16, SRK_CSR_FLAGS=$fl"; SRK_CSR_FLAGS=$fl timeout -k 10 300 python bench.py --mode csr16 --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2_bench_csr16_fl$fl.json 2> gpurun_out/r2_bench_csr16_fl$fl.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_csr16_fl$fl.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], json.dumps(d["kernels"]))
PY
tail -2 gpurun_out/r2_bench_csr16_fl$fl.err
done
export SRK_REAL_CFG5=1
export SRK_SWEEP="SRK_CSR_FLAGS=8;SRK_CSR_FLAGS=0"
timeout -k 10 400 python scripts/csr_shape_bench.py cfg5_s1_final cfg4_n8_final 2>&1 | tee gpurun_out/r2_csr_shapes_prefetch.jsonl | cut -c1-120
