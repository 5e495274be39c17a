#!/usr/bin/env python
"""Condense an Nsight Compute report (.ncu-rep, read with `ncu -i`) into the JSON kept under
profiles/: per captured kernel the metrics DESIGN.md quotes (duration, tensor-pipe activity, DRAM and
L2 traffic, clocks) and the stall reasons of the hottest SASS instructions.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_full_<what>.json
"""
import collections
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
    "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "l1tex__t_bytes.sum.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def raw_page(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        k = {"kernel": d[hdr.index("Kernel Name")]}
        for m in METRICS:
            for i, h in enumerate(hdr):
                if h == m or h.endswith("." + m):
                    k[m] = [d[i], units[i]]
                    break
        out.append(k)
    return out


def stall_page(rep, top=12):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    out = []
    for n, s in enumerate(starts):
        end = starts[n + 1] if n + 1 < len(starts) else len(rows)
        hdr, data = rows[s + 1], rows[s + 2:end]
        if "# Samples" not in hdr:
            continue
        ix = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        total = sum(int(r[ix["# Samples"]]) for r in data)
        if not total:
            continue
        agg = collections.Counter()
        for r in data:
            for h in stalls:
                agg[h] += int(r[ix[h]])
        hot = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]
        if any(o["kernel"] == rows[s][1] and o["samples"] == total for o in out):
            continue                      # the source page lists every function twice
        out.append({"kernel": rows[s][1], "samples": total,
                    "stall_share": {k: round(v / total, 4) for k, v in agg.most_common(8)},
                    "hottest": [{"sass": r[1].strip(), "samples": int(r[ix["# Samples"]]),
                                 "executed": int(r[ix["Instructions Executed"]])} for r in hot]})
    return out


def traffic_entries(kernels):
    """{bench kernel name: bytes per launch} for the kernels bench.py's roofline can name."""
    names = {"i8x2_kernel<2, 0>": ("x2_half_mid", 2), "i8x2_kernel<2, 1>": ("x2_half_final", 2),
             "csr_gather_kernel<unsigned short, 512, 0>": ("csr16_half_first", 16),
             "csr_gather_kernel<unsigned short, 512, 2>": ("csr16_half_final", 16),
             "i8x2_kernel<3, 0>": ("x2_half_mid", 3), "i8x2_kernel<3, 1>": ("x2_half_final", 3)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    out = {}
    for k in kernels:
        for pat, (name, ns) in names.items():
            if pat in k["kernel"] and "dram__bytes_read.sum" in k:
                b = sum(float(k[m][0]) * scale[k[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                out[(name, ns)] = b
    return out


if __name__ == "__main__":
    rep, dst = sys.argv[1], sys.argv[2]
    kernels = raw_page(rep)
    json.dump({"report": rep, "kernels": kernels, "stalls": stall_page(rep)}, open(dst, "w"), indent=1)
    print(dst)
    if len(sys.argv) > 3:                 # ncu_traffic.json  "n32768_g1": merge the DRAM bytes per launch
        path, tag = sys.argv[3], sys.argv[4]
        try:
            cur = json.load(open(path))
        except FileNotFoundError:
            cur = {}
        for (name, ns), b in traffic_entries(kernels).items():
            cur.setdefault(name, {})[f"ns{ns}_{tag}"] = b
        pipe = "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"
        names = {"i8x2_kernel<2, 0>": "x2_half_mid", "i8x2_kernel<2, 1>": "x2_half_final"}
        for k in kernels:
            for pat, name in names.items():
                if pat in k["kernel"] and pipe in k:
                    cur.setdefault("_tensor_pipe_active", {}).setdefault(name, {})[f"ns2_{tag}"] = float(k[pipe][0]) / 100.0
        cur["_source"] = cur.get("_source", {})
        cur["_source"][tag] = dst
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
        print(path)
