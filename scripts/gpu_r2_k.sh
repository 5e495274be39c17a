#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== csr gather kernel tests + logical shards"; timeout -k 5 400 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_kernels.py tests/test_gpu_local_cluster.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -8 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
echo "== bench csr16"; timeout -k 10 300 python bench.py --mode csr16 --steps 5 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2_bench_csr16.json 2> gpurun_out/r2_bench_csr16.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_csr16.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["kernels"], d.get("parity"))
PY
bash scripts/gpu_r2_ab.sh
