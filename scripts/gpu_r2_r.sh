#!/bin/bash
# one rank's cfg5 launches replayed with the bench's own graph, split settings swept; then the N=1 bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export SRK_REAL_CFG5=1
export SRK_SWEEP="SRK_SPLIT_MIN=0;SRK_SPLIT_MIN=256,SRK_SPLIT_PIECE=512;SRK_SPLIT_MIN=2048,SRK_SPLIT_PIECE=1024;SRK_SPLIT_MIN=64,SRK_SPLIT_PIECE=256;SRK_SPLIT_MIN=1024,SRK_SPLIT_PIECE=256"
timeout -k 10 700 python scripts/csr_shape_bench.py cfg5_s2_first cfg5_s2_final cfg5_s1_final cfg5_s1_first 2>&1 | tee gpurun_out/r2_csr_split_shapes_real.jsonl | cut -c1-330
unset SRK_SWEEP
echo "== bench N=1"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_k20.json 2> gpurun_out/r2_bench_n1_k20.err
echo "rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1_k20.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], json.dumps(d["e2e"]), json.dumps(d["csr_path"]["kernels"]), d["csr_path"]["value"], json.dumps(d.get("cpu_baseline")))
PY
tail -3 gpurun_out/r2_bench_n1_k20.err
