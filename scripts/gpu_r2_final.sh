#!/bin/bash
# Final check of the round on one GPU: whole GPU suite, smoke(), the default bench line at the driver's K / W
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu suite"; timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?"; tail -5 gpurun_out/r2_gpu_tests.log
echo "== smoke"; timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench default"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_k20.json 2> gpurun_out/r2_bench_n1_k20.err
echo "rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1_k20.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], json.dumps(d["kernels"]), d["clocks"], d["roofline"]["frac"], d["roofline"].get("frac_int8"))
print(json.dumps(d["e2e"]))
print(d["csr_path"]["value"], json.dumps(d["csr_path"]["kernels"]), d["csr_path"]["parity"]["max_abs"], d["parity"]["max_abs"])
print(json.dumps(d["cpu_baseline"]), d["gpu_launches"])
PY
tail -3 gpurun_out/r2_bench_n1_k20.err
