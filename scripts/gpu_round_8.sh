#!/bin/bash
# 8-GPU session (gpurun --gpus 8): headline bench at N=8 and BASELINE cfg5 at full size.
# usage: gpu_round_8.sh [bench] [cfg5]   (default: both)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
what="${*:-bench cfg5}"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [[ "$what" == *bench* ]]; then
echo "== bench N=8"; timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
echo "rc=$?"; grep '^{' gpurun_out/scale_n8.json | cut -c1-1800; grep -v "OMP_NUM\|\*\*\*" gpurun_out/scale_n8.err | tail -5 | cut -c1-300
fi
if [[ "$what" == *cfg5* ]]; then
echo "== cfg5 full size, K=3"; timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 scripts/cfg5_sharded.py 1.0 3 > gpurun_out/cfg5_full.json 2> gpurun_out/cfg5_full.err
echo "rc=$?"; grep '^{' gpurun_out/cfg5_full.json | cut -c1-3000; grep -v "OMP_NUM\|\*\*\*" gpurun_out/cfg5_full.err | tail -15 | cut -c1-300
fi
