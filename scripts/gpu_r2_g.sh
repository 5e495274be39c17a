#!/bin/bash
# Round 2: int8 tensor peak microbenchmark, CSR kernel tests + bench after the triangular grid /
# symmetric quantiser changes, ncu of csr16.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== i8 peak"; timeout -k 5 120 scripts/micro/i8_peak > gpurun_out/r2_micro_i8_peak.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2_micro_i8_peak.txt
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
echo "== csr gather kernel tests"; timeout -k 5 180 python -m pytest tests/test_gpu_csr_gather.py tests/test_gpu_kernels.py -x -q > gpurun_out/r2_csr_tests.log 2>&1
rc=$?; tail -8 gpurun_out/r2_csr_tests.log; echo "csr tests rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
echo "== logical shards + parity"; timeout -k 10 900 python -m pytest tests/test_gpu_local_cluster.py tests/test_gpu_parity.py -x -q > gpurun_out/r2_parity_tests.log 2>&1
echo "rc=$?"; tail -8 gpurun_out/r2_parity_tests.log
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
echo "== ncu full csr16"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"csr_gather|quantize|row_unit" -s 4 -c 4 -o gpurun_out/r2_prof_csr16 -f python bench.py --mode csr16 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_ncu_csr16.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_ncu_csr16.log
