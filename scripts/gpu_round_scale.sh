#!/bin/bash
# Scaling session on an 8-GPU box: sharded parity test at 8, bench at 8, 4, 2, 1.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== dist test (8)"; timeout -k 10 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/dist_tests_n8.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/dist_tests_n8.log
for N in 8 4 2; do
  echo "== bench N=$N"; timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  echo "rc=$?"; grep '^{' gpurun_out/scale_n$N.json | python -c "
import json,sys
for line in sys.stdin:
    l=json.loads(line); print(l['n_gpus'], round(l['value'],2), 'it/s', round(l['ms_per_step'],2), 'ms', {k:round(v['ms'],2) for k,v in l['kernels'].items()}, l['clocks'], l.get('e2e',{}).get('value'))"
  tail -3 gpurun_out/scale_n$N.err | cut -c1-300
done
echo "== bench N=1"; timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
grep '^{' gpurun_out/scale_n1.json | python -c "
import json,sys
for line in sys.stdin:
    l=json.loads(line); print(l['n_gpus'], round(l['value'],2), 'it/s', round(l['ms_per_step'],2), 'ms', {k:round(v['ms'],2) for k,v in l['kernels'].items()}, l['clocks'], l.get('e2e',{}).get('value'))"
