#!/bin/bash
# One GPU session: unit tests of the paired-SM kernel first (short timeout: a hang must not eat
# the box), then the whole GPU suite, full-size parity, the bench in its variants and ncu captures.
# usage: gpu_round.sh [quick]
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== x2 unit tests"; timeout -k 5 150 python -m pytest tests/test_gpu_x2.py -x -q > gpurun_out/x2_tests.log 2>&1
rc=$?; tail -15 gpurun_out/x2_tests.log; echo "x2 tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
echo "== full gpu suite"; timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
echo "gpu suite rc=$?"; tail -15 gpurun_out/gpu_tests.log
echo "== parity n=32768"; timeout -k 10 600 python scripts/parity_full.py > gpurun_out/parity_full.json 2> gpurun_out/parity_full.err
echo "rc=$?"; cat gpurun_out/parity_full.json | cut -c1-1500; tail -3 gpurun_out/parity_full.err
echo "== bench x2 auto"; timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_x2_auto.json 2> gpurun_out/bench_x2_auto.err
echo "rc=$?"; cat gpurun_out/bench_x2_auto.json; tail -3 gpurun_out/bench_x2_auto.err
echo "== bench x2 ns3"; timeout -k 10 600 python bench.py --steps 5 --warmup 3 --slices 3 --no-e2e --no-cpu > gpurun_out/bench_x2_ns3.json 2> gpurun_out/bench_x2_ns3.err
echo "rc=$?"; cat gpurun_out/bench_x2_ns3.json
if [ "${1:-}" = "quick" ]; then exit 0; fi
echo "== ncu launch list"; timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_x2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"
echo "== ncu full"; timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:i8x2 -s 3 -c 2 -o gpurun_out/prof_x2 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; ls -la gpurun_out
