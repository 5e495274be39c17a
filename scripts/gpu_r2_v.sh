#!/bin/bash
# Round 2 checkpoint on one GPU: whole GPU suite, compute-sanitizer over the small cases of the CSR gather
# kernels (incl. ACCUM / FINISH), the default bench line at the driver's K / W, ncu --set full of one rank's
# cfg5 S1 second half (ACCUM + FINISH) and S2 first half (ACCUM + FIRST).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu suite"; timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1
echo "gpu suite rc=$?"; tail -6 gpurun_out/r2_gpu_tests.log
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool (CSR gather kernels)"
  timeout -k 10 420 compute-sanitizer --tool $tool --error-exitcode 9 \
    python -m pytest tests/test_gpu_csr_gather.py -x -q \
    -k "split or finish or accum or (first_half and 37) or (final_symmetric and 70) or (final_transposed and 70) or quantize" \
    > gpurun_out/r2_sanitize_csr_$tool.log 2>&1
  echo "rc=$?"; tail -6 gpurun_out/r2_sanitize_csr_$tool.log
done
echo "== bench default"; timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_k20.json 2> gpurun_out/r2_bench_n1_k20.err
echo "rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1_k20.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], json.dumps(d["kernels"]), d["clocks"])
print(json.dumps(d["e2e"]))
print(d["csr_path"]["value"], json.dumps(d["csr_path"]["kernels"]), d["csr_path"]["parity"]["max_abs"], d["parity"]["max_abs"])
PY
tail -3 gpurun_out/r2_bench_n1_k20.err
echo "== ncu full, cfg5 shapes of one rank"
SRK_REAL_CFG5=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"csr_gather|csr_finish" -c 4 -o gpurun_out/r2_prof_cfg5_shapes -f python scripts/csr_shape_bench.py cfg5_s1_final cfg5_s2_first > gpurun_out/r2_ncu_cfg5_shapes.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_ncu_cfg5_shapes.log | cut -c1-300
ls -la gpurun_out/r2_prof_cfg5_shapes.ncu-rep
