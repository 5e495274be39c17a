#!/bin/bash
# N GPUs (8): BASELINE cfg5 at full size on the CSR SpMM path, then the cfg4 bench line at the driver's K / W
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
tr="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
echo "== cfg5 scale=1 mode=csr16 N=$n"; timeout -k 10 900 $tr --master-port 29542 bench.py --config cfg5 --scale 1 --mode csr16 --gpus $n --steps 3 --warmup 3 > gpurun_out/r2_cfg5_csr16_n$n.json 2> gpurun_out/r2_cfg5_csr16_n$n.err
echo "rc=$?"; grep '^{' gpurun_out/r2_cfg5_csr16_n$n.json | cut -c1-3500; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2_cfg5_csr16_n$n.err | tail -8 | cut -c1-400
echo "== bench N=$n (cfg4)"; timeout -k 10 600 $tr --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale_n${n}_k20.json 2> gpurun_out/r2_scale_n${n}_k20.err
echo "rc=$?"; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_scale_n${n}_k20.json").read().strip().splitlines() if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], json.dumps(d["kernels"]), json.dumps(d.get("parity"))[:300])
print(json.dumps(d["e2e"]))
print(d["csr_path"]["value"], json.dumps(d["csr_path"]["kernels"]), json.dumps(d["csr_path"].get("parity"))[:300])
PY
grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2_scale_n${n}_k20.err | tail -5 | cut -c1-300
