#!/bin/bash
# 8 GPUs: BASELINE cfg5 at full size on the CSR SpMM path (final state of the round)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
tr="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
echo "== cfg5 scale=1 mode=csr16 N=$n"; timeout -k 10 600 $tr --master-port 29542 bench.py --config cfg5 --scale 1 --mode csr16 --gpus $n --steps 3 --warmup 3 > gpurun_out/r2_cfg5_csr16_n$n.json 2> gpurun_out/r2_cfg5_csr16_n$n.err
echo "rc=$?"; grep '^{' gpurun_out/r2_cfg5_csr16_n$n.json | cut -c1-3500; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2_cfg5_csr16_n$n.err | tail -8 | cut -c1-400
