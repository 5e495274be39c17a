#!/bin/bash
# Per-tile timestamps of the paired-SM kernel (SRK_X2_TRACE): mainloop start/end and epilogue
# start/end of every CTA pair; summarise with scripts/trace_report.py.  Extra arguments go to bench.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/trace
SRK_X2_TRACE=gpurun_out/trace/x2 timeout -k 5 300 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu "$@" 2>&1 | cut -c1-100
(cd gpurun_out/trace && ls | head -n -2 | xargs rm -f; ls)
