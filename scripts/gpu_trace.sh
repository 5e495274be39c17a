#!/bin/bash
# Per-tile timestamps of the paired-SM kernel (SRK_X2_TRACE): mainloop start/end and epilogue
# start/end of every CTA pair, with the lockstep on (default) and off; summarise with
# scripts/trace_report.py.  Extra arguments go to bench.py.
cd "$(dirname "$0")/.."
for mode in on off; do
  mkdir -p gpurun_out/trace_$mode
  if [ $mode = off ]; then export SRK_X2_LOCKSTEP=0; fi
  SRK_X2_TRACE=gpurun_out/trace_$mode/x2 timeout -k 5 300 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu "$@" 2>&1 | cut -c1-100
  (cd gpurun_out/trace_$mode && ls | head -n -2 | xargs rm -f; ls)
done
