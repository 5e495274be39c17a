#!/bin/bash
# A/B of the dense kernel's L2 policy / band height (1 GPU): ms per iteration, per-kernel times, clocks.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for variant in "0:8" "4:8" "8:8" "4:4" "8:4" "4:12" "8:16" "0:8"; do
  flags=${variant%%:*}; group=${variant##*:}
  SRK_X2_FLAGS=$flags SRK_X2_GROUP=$group timeout -k 10 200 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-csr --no-parity > gpurun_out/ab_${flags}_${group}.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${flags}_${group}.json").read().strip().splitlines()[-1])
    print("flags=$flags group=$group", round(d["ms_per_step"],2), {k: round(v["ms"],2) for k,v in d["kernels"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("flags=$flags group=$group failed", e); print(open("gpurun_out/ab.err").read()[-800:])
PY
done
