#!/bin/bash
# single-GPU csr16 halves as ACCUM + streaming passes: kernel tests, parity tests, bench of the csr16 path both ways,
# FINISH with 3 / 4 CTAs per SM on the cfg5 S1 shape
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -5 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
for via in 0 1; do
echo "== bench csr16, SRK_CSR_VIA_ACCUM=$via"; SRK_CSR_VIA_ACCUM=$via timeout -k 10 300 python bench.py --mode csr16 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench_csr16_via$via.json 2> gpurun_out/r2_bench_csr16_via$via.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_csr16_via$via.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], json.dumps(d["kernels"]), d.get("parity",{}).get("max_abs"))
PY
tail -2 gpurun_out/r2_bench_csr16_via$via.err
done
export SRK_REAL_CFG5=1
export SRK_SWEEP="SRK_CSR_FLAGS=0;SRK_CSR_FLAGS=4"
timeout -k 10 400 python scripts/csr_shape_bench.py cfg5_s1_final cfg4_n8_final 2>&1 | tee gpurun_out/r2_csr_shapes_finish_ctas.jsonl | cut -c1-200
