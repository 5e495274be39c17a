#!/bin/bash
# N GPUs: short cfg4 bench line with parity (final state of the row-sharded paths)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
tr="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout -k 10 200 $tr --master-port 29541 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_scale_n${n}_final.json 2> gpurun_out/r2_scale_n${n}_final.err
echo "rc=$?"; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_scale_n${n}_final.json").read().strip().splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["parity"]["max_abs"])
print(json.dumps(d["e2e"])[:400])
print(d["csr_path"]["value"], json.dumps(d["csr_path"]["kernels"]), d["csr_path"]["parity"]["max_abs"])
PY
grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2_scale_n${n}_final.err | tail -4 | cut -c1-300
