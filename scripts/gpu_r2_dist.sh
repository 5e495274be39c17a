#!/bin/bash
# Round 2, multi-GPU session (gpurun --gpus 2|4|8): parity of the row-sharded solvers against the oracle,
# the headline bench with its parity object, and BASELINE cfg5 (smoke scale unless CFG5_SCALE=1).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l)
scale=${CFG5_SCALE:-0.125}
tr="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== row-sharded parity on $n GPUs"
timeout -k 10 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_zz_dist_csr.py -x -q -k "row_sharded" > gpurun_out/r2_dist_tests_n$n.log 2>&1
echo "rc=$?"; tail -12 gpurun_out/r2_dist_tests_n$n.log
fi
echo "== bench N=$n (cfg4)"; timeout -k 10 600 $tr --master-port 29541 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
echo "rc=$?"; grep '^{' gpurun_out/r2_scale_n$n.json | cut -c1-2500; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2_scale_n$n.err | tail -5 | cut -c1-300
for mode in ${CFG5_MODES:-csr16 i8}; do
echo "== cfg5 scale=$scale mode=$mode N=$n"; timeout -k 10 900 $tr --master-port 29542 bench.py --config cfg5 --scale $scale --mode $mode --gpus $n --steps 3 --warmup 3 > gpurun_out/r2_cfg5_${mode}_n$n.json 2> gpurun_out/r2_cfg5_${mode}_n$n.err
echo "rc=$?"; grep '^{' gpurun_out/r2_cfg5_${mode}_n$n.json | cut -c1-3500; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2_cfg5_${mode}_n$n.err | tail -8 | cut -c1-400
done
if [ "${REF_ARM:-0}" = "1" ]; then
echo "== reference arm under torchrun"; timeout -k 10 600 $tr --master-port 29543 bench.py --impl reference --gpus $n --steps 2 --warmup 1 > gpurun_out/r2_ref_n$n.json 2> gpurun_out/r2_ref_n$n.err
echo "rc=$?"; grep '^{' gpurun_out/r2_ref_n$n.json | cut -c1-800
fi
