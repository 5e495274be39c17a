#!/bin/bash
# A/B experiments on the paired-SM kernel through SRK_X2_DEBUG (timing only for bits 1/2/4, which
# switch parts of the FINAL epilogue off; bits 8/16 only change L2 eviction hints, results stay exact):
#   8 = epilogue loads/stores with the default policy instead of evict_first
#  16 = TMA operand loads with the default policy instead of evict_last
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== x2 unit tests"; timeout -k 5 200 python -m pytest tests/test_gpu_x2.py -x -q 2>&1 | tail -3
for d in ${@:-0 8 16 24}; do
  echo "== SRK_X2_DEBUG=$d"
  SRK_X2_DEBUG=$d timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print({k:round(v['ms'],2) for k,v in l['kernels'].items()}, round(l['ms_per_step'],2), l['clocks'])"
done
