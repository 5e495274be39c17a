#!/bin/bash
# Multi-GPU session (run with gpurun --gpus N): sharded parity test, bench at N and at 1.
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== x2 unit tests (quick)"; timeout -k 5 150 python -m pytest tests/test_gpu_x2.py -x -q > gpurun_out/x2_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/x2_tests.log
echo "== dist test"; timeout -k 10 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/dist_tests_n$N.log 2>&1
rc=$?; echo "rc=$rc"; tail -30 gpurun_out/dist_tests_n$N.log
if [ $rc -ne 0 ]; then
  echo "== dist test, staged exchange"; SIMRANK_B200_EXCHANGE=staged timeout -k 10 900 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/dist_tests_staged_n$N.log 2>&1
  echo "rc=$?"; tail -30 gpurun_out/dist_tests_staged_n$N.log
fi
echo "== bench N=$N (peer)"; timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
echo "== bench N=$N (staged)"; SIMRANK_B200_EXCHANGE=staged timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_staged_n$N.json 2> gpurun_out/bench_staged_n$N.err
echo "rc=$?"; cat gpurun_out/bench_staged_n$N.json; tail -5 gpurun_out/bench_staged_n$N.err
echo "== bench N=1"; timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
