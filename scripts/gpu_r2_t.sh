#!/bin/bash
# N GPUs: the cfg4 bench line (dense chain + sharded csr16 csr_path + e2e through the default 'auto' mode)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
tr="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
echo "== bench N=$n (cfg4)"; timeout -k 10 600 $tr --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 --no-parity > gpurun_out/r2_scale_n${n}_k20.json 2> gpurun_out/r2_scale_n${n}_k20.err
echo "rc=$?"; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_scale_n${n}_k20.json").read().strip().splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], json.dumps(d["kernels"]))
print(json.dumps(d["e2e"]))
print(d["csr_path"]["value"], json.dumps(d["csr_path"]["kernels"]))
PY
grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2_scale_n${n}_k20.err | tail -5 | cut -c1-300
