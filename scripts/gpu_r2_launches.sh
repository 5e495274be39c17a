#!/bin/bash
# ncu launch list (per-launch durations, cold-cache and serialised) of the default bench command's kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_n1.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-parity --no-e2e > gpurun_out/r2_launches_bench.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_bench_n1.csv")) if len(r)>10]
hdr=rows[0]; k=hdr.index("Kernel Name"); v=hdr.index("Metric Value"); u=hdr.index("Metric Unit")
agg=collections.defaultdict(list)
for r in rows[1:]:
    val=float(r[v].replace(",","")); unit=r[u]
    ms = val/1e6 if unit in ("ns","nsecond") else val/1e3 if unit in ("us","usecond") else val
    agg[r[k][:70]].append(ms)
for name,t in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:16]:
    print(f"{len(t):4d} x {sum(t)/len(t):9.3f} ms  {name}")
PY
