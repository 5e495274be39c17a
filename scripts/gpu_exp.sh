#!/bin/bash
# A/B of the lockstep throttle of the paired-SM kernel: SRK_X2_LOCKSTEP=0 lets the CTA pairs run
# free, "units,lag" changes the granularity.  usage: gpu_exp.sh SLICES [settings...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ns=${1:-auto}; shift
for d in ${@:-8,1 0}; do
  echo "== slices=$ns SRK_X2_LOCKSTEP=$d"
  SRK_X2_LOCKSTEP=$d timeout -k 5 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --slices $ns 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print({k:round(v['ms'],2) for k,v in l['kernels'].items()}, round(l['ms_per_step'],2), l['clocks'])"
done
