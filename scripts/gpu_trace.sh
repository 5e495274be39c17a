#!/bin/bash
# Per-tile timestamps of the paired-SM kernel (SRK_X2_TRACE): mainloop start/end and epilogue
# start/end of every CTA pair, for each SRK_X2_DEBUG variant given on the command line.
cd "$(dirname "$0")/.."
for d in ${@:-0}; do
  mkdir -p gpurun_out/trace_d$d
  SRK_X2_DEBUG=$d SRK_X2_TRACE=gpurun_out/trace_d$d/x2 timeout -k 5 300 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu 2>&1 | cut -c1-100
  (cd gpurun_out/trace_d$d && ls | head -n -2 | xargs rm -f; ls)
done
