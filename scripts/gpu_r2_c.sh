#!/bin/bash
# Round 2, CSR gather kernel session (1 GPU): kernel tests first under a short timeout (a hang must
# not eat the box), then parity, then the bench in the CSR modes with A/B flags, then ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== csr gather kernel tests"; timeout -k 5 180 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_kernels.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -25 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
echo "== logical shards on one GPU"; timeout -k 10 600 python -m pytest tests/test_gpu_local_cluster.py -x -q > gpurun_out/r2_local_cluster_tests.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/r2_local_cluster_tests.log
echo "== parity suite"; timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2_parity_tests.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/r2_parity_tests.log
for variant in "csr16:0" "csr:0"; do
  mode=${variant%%:*}; flags=${variant##*:}
  echo "== bench $mode flags=$flags"
  SRK_CSR_FLAGS=$flags timeout -k 10 300 python bench.py --mode $mode --steps 5 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2_bench_${mode}_f$flags.json 2> gpurun_out/r2_bench_${mode}_f$flags.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${mode}_f$flags.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["kernels"], d.get("parity"), d["clocks"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_bench_${mode}_f$flags.err").read()[-1500:])
PY
done
echo "== ncu full csr16"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:csr_gather -s 2 -c 2 -o gpurun_out/r2_prof_csr16 -f python bench.py --mode csr16 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_ncu_csr16.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_ncu_csr16.log
