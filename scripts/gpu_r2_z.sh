#!/bin/bash
# streaming passes with cp.async-staged tiles: kernel / parity / logical-shard tests, csr16 bench, one rank's shapes
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_local_cluster.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -3 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
echo "== bench csr16"; timeout -k 10 300 python bench.py --mode csr16 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench_csr16.json 2> gpurun_out/r2_bench_csr16.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_csr16.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], json.dumps(d["kernels"]), d.get("parity",{}).get("max_abs"))
PY
tail -2 gpurun_out/r2_bench_csr16.err
export SRK_REAL_CFG5=1
timeout -k 10 400 python scripts/csr_shape_bench.py cfg5_s1_final cfg4_n8_final cfg5_s2_final 2>&1 | tee gpurun_out/r2_csr_shapes_fastpath.jsonl | cut -c1-120
