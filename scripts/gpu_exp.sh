#!/bin/bash
# Profiling experiments on the FINAL epilogue (SRK_X2_DEBUG switches parts of it off; results are
# then wrong on purpose -- timing only).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 0 1 2 4 3 7; do
  echo "== SRK_X2_DEBUG=$d"
  SRK_X2_DEBUG=$d timeout -k 5 200 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print({k:round(v['ms'],2) for k,v in l['kernels'].items()}, round(l['ms_per_step'],2), l['clocks'])"
done
