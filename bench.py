#!/usr/bin/env python
"""Headline benchmark: SimRank iterations/sec at n = 32768 (BASELINE.json configs[3], "cfg4").

    python bench.py --gpus N --steps K --warmup W            this engine (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    the reference's CPU path (numpy f64)

A step is ONE iteration of ``S <- C * G S G^T; diag <- 1`` with the fused max|dS| reduction and
its read-back (SimRank.py:130-140), on the synthetic dense-regime directed graph of
simrank_b200/synth.py (n = 32768, m = 2097152, seed 4).  Prints one JSON line (see README /
DESIGN.md "Measurement" for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

if "--impl" in sys.argv and "reference" in sys.argv:
    # The reference arm is numpy/OpenBLAS on the host cores.  torchrun exports OMP_NUM_THREADS=1 to its
    # workers, which would time the reference on ONE core: give rank 0 (the only rank that works in
    # this arm) the whole host back, before numpy loads its BLAS.
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "simrank_iterations_per_sec_n32768"
UNIT = "iter/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--n", type=int, default=32768, help="nodes (default: the BASELINE cfg4 size)")
    ap.add_argument("--mean-degree", type=int, default=64)
    ap.add_argument("--mode", default="i8", choices=["i8", "csr", "csr16"],
                    help="dense tensor-core chain (graded), the float64 CSR SpMM path, or the fixed-point CSR SpMM path")
    ap.add_argument("--slices", default="auto", help="uint8 planes per matrix: 2, 3, 4 or auto (error-bound driven)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the float64 re-run that checks the timed solver's result (outside the timed region)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload(args):
    from simrank_b200 import graph, synth
    n, m = args.n, args.n * args.mean_degree
    frm, to = synth.directed_edges(n, m, 0.5, 4)
    op = graph.operator_from_edges(to, frm, n, n)            # G[to, from] = 1/indeg(to)
    name = (f"cfg4: SimRank on a synthetic dense-regime directed graph, n={n}, m={m} unique edges "
            f"(power-law alpha=0.5, seed 4), unweighted, C=0.8")
    return op, (frm, to), name


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="B200_PROFILING.md fallback")


def ncu_pipe_active(kernel, ns, n, world):
    """Fraction of cycles the tensor pipe was active in that capture (ncu
    sm__pipe_tensor_cycles_active_realtime), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get("_tensor_pipe_active", {}).get(kernel, {}).get(f"ns{ns}_n{n}_g{world}")


def ncu_traffic(kernel, ns, n, world):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of ``kernel`` from the
    committed `ncu --set full` capture of this very workload (profiles/ncu_traffic.json, written by
    scripts/ncu_summary.py); None when no capture matches."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(kernel, {}).get(f"ns{ns}_n{n}_g{world}")


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self._stop = index, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import psutil
    from oracle import cpu_baseline
    try:                                        # the BLAS pool may have been created with one thread already
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1, user_api="blas")
    except Exception:
        pass
    op, _, name = workload(args)
    res = cpu_baseline.iterations_per_second(op.indptr, op.indices, op.g, op.M, target_seconds=args.cpu_seconds,
                                             steps=args.steps, warmup=args.warmup,
                                             max_bytes=int(psutil.virtual_memory().available * 0.8))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / res["value"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name},
            "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                             "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- engine arm
def run_engine(args):
    args.slices = None if args.slices in ("auto", None) else int(args.slices)
    import torch
    import torch.distributed as dist
    from simrank_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = engine.require_cuda(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    op, edges, name = workload(args)
    n = op.M

    if world > 1:
        from simrank_b200 import dist as sdist
        solver = sdist.ShardedDirectedSolver(op, 0.8, mode=args.mode, ns=args.slices, device=dev)
        halves = solver.halves
    else:
        dop = engine.DeviceOperator(op, dev)
        solver = engine.DirectedSolver(dop, 0.8, mode=args.mode, ns=args.slices)
        halves = [solver.half]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        solver.step()
    for h in halves:
        h.events = []
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from simrank_b200 import _lib as _srk_lib
    launches_before = _srk_lib.LAUNCHES
    with ClockSampler(local) as clocks:
        t0.record()
        last = None
        for _ in range(args.steps):
            last = solver.step()
        t1.record()
        sync_all()
    launches = _srk_lib.LAUNCHES - launches_before          # kernels of libsimrank_b200 enqueued by the timed steps (this rank)
    ms = t0.elapsed_time(t1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # per-kernel durations from the events recorded inside the timed steps
    per = {}
    for h in halves:
        for nm, a, b in h.events:
            per.setdefault(nm, []).append(a.elapsed_time(b))
        h.events = None
    pk = peaks()
    flops_half = 2.0 * n * n * n / world                        # algorithmic: one n x n x n product, row-sharded
    kernels = {k: {"ms": statistics.mean(v), "launches": len(v)} for k, v in per.items()}
    used = sorted(set(x for h in halves for x in getattr(h, "slices_used", [])[-args.steps:])) or [args.slices or 3]
    if args.mode == "i8":
        gemms = {k: v for k, v in kernels.items() if "half" in k}
        dom = max(gemms, key=lambda k: gemms[k]["ms"])
        ach = flops_half / (kernels[dom]["ms"] * 1e-3) / 1e12
        sym = args.mode == "i8" and world == 1 and dom.endswith("final")
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor_sustained"], "traffic": ncu_traffic(dom, used[-1], n, world),
                "peak_kind": f"dense bf16 sustained, {pk['source']}; burst {pk['tensor_burst']}",
                "executed_int8_tops": ach * used[-1] * (0.5 if sym else 1.0),
                "tensor_pipe_active_ncu": ncu_pipe_active(dom, used[-1], n, world),
                "note": ("achieved = algorithmic 2n^3 flop of one half-product / mean launch time inside the timed "
                         f"steps; the kernel executes {used} u8 x u8 -> s32 tcgen05 products per algorithmic one"
                         + ("; this launch computes only the upper triangle of the symmetric result" if sym else ""))}
    else:
        halves_only = {k: v for k, v in kernels.items() if "half" in k}
        dom = max(halves_only, key=lambda k: halves_only[k]["ms"])
        by = (3 if dom.endswith("final") else 2) * n * n * 8.0 / world
        ach = by / (kernels[dom]["ms"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": ncu_traffic(dom, 0, n, world), "peak_kind": pk["source"],
                "note": "algorithmic bytes (SURVEY.md 8d): first half 2 n^2 s, second half 3 n^2 s (s = 8)",
                "iteration": {"algorithmic_bytes": 5.0 * n * n * 8.0 / world,
                              "achieved": 5.0 * n * n * 8.0 / world / (ms * 1e-3) / 1e9,
                              "frac": 5.0 * n * n * 8.0 / world / (ms * 1e-3) / 1e9 / pk["hbm"]}}

    line = {"metric": METRIC, "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": {"i8": f"u8 x u8 -> s32 ({used} fixed-point planes), f64 epilogue", "csr": "f64",
                      "csr16": "u16 gather -> exact u32 sums, f64 epilogue"}[args.mode],
            "data": "synthetic",
            "config": {"workload": name, "mode": args.mode, "slices": args.slices or "auto", "slices_used": used,
                       "l2": "operands (>= 1 GB) are far larger than the 126 MB L2; no flush needed",
                       "parallelism": f"S row-sharded over {world} GPU(s)"},
            "algorithmic_tflops": 4.0 * n ** 3 / (ms * 1e-3) / 1e12,
            "roofline": roof, "kernels": kernels, "gpu_launches": launches,
            "last_maxdiff": last if not isinstance(last, tuple) else list(last)}

    if not args.no_parity:
        line["parity"] = run_parity(args, solver, halves, op, dev, world, args.warmup + args.steps)

    clk = clocks.summary()
    if world > 1:
        # the per-kernel times above are rank 0's; the spread over the ranks shows who waits for whom
        every = [None] * world
        dist.all_gather_object(every, {k: v["ms"] for k, v in kernels.items()})
        line["kernels_rank_spread_ms"] = {k: [round(min(e.get(k, 0.0) for e in every), 3),
                                              round(max(e.get(k, 0.0) for e in every), 3)] for k in kernels}
        gathered = [None] * world
        dist.all_gather_object(gathered, clk)
        mhz = [g["sm_mhz"] for g in gathered if g["sm_mhz"]]
        clk = {"sm_mhz": min(mhz) if mhz else None, "sm_max_mhz": clk["sm_max_mhz"],
               "reasons": sorted(set(r for g in gathered for r in g["reasons"]))}
    line["clocks"] = clk

    del solver, halves
    torch.cuda.empty_cache()
    # ---- end to end through the public API: DataFrame in -> DataFrame out, host buffers
    if not args.no_e2e:
        line["e2e"] = run_e2e(args, edges, n, world, rank)
        torch.cuda.empty_cache()

    # ---- the reference's CPU path on this host's cores (rank 0, N = 1 only)
    if world == 1 and not args.no_cpu:
        import psutil
        from oracle import cpu_baseline
        res = cpu_baseline.iterations_per_second(op.indptr, op.indices, op.g, n, target_seconds=args.cpu_seconds,
                                                 max_bytes=int(psutil.virtual_memory().available * 0.8))
        line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                                "sample": res["sample"]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


PARITY_TOL = 1e-6          # north_star: max-abs 1e-6 against the float64 reference after K iterations


def run_parity(args, solver, halves, op, dev, world, iterations):
    """Check the matrix the timed solver ended with -- OUTSIDE the timed region.  Every rank re-runs
    the same ``iterations`` updates from S = I on its own GPU with the single-GPU float64 CSR path
    (the exact-arithmetic mode of this engine, itself held to 1e-12 against the numpy oracle by
    tests/test_gpu_parity.py) and compares ALL rows it owns (the whole matrix at N = 1).  Also checks
    the unit diagonal.  Raises (non-zero exit) when the deviation exceeds PARITY_TOL."""
    import torch
    import torch.distributed as dist
    from simrank_b200 import engine
    h = halves[0]
    rows, row0 = (h.rows, h.row0) if world > 1 else (op.M, 0)
    mine = h.S[:rows, :op.M]
    ref = engine.DirectedSolver(engine.DeviceOperator(op, dev), 0.8, mode="csr")
    for _ in range(iterations):
        ref.step()
    want = ref.S[row0:row0 + rows]
    worst = float((mine - want).abs().max().item()) if rows else 0.0
    diag_ok = bool((mine[torch.arange(rows, device=dev), row0 + torch.arange(rows, device=dev)] == 1.0).all().item()) \
        if rows else True
    top_ok = None
    if rows:                                   # node-index parity: top-10 of 64 sampled rows, off-diagonal
        pick = torch.linspace(0, rows - 1, min(rows, 64), device=dev).long()
        a, b = mine[pick].clone(), want[pick].clone()
        a[torch.arange(len(pick), device=dev), row0 + pick] = -1.0
        b[torch.arange(len(pick), device=dev), row0 + pick] = -1.0
        ia, va = engine.topk_rows(a.contiguous(), 10)
        ib, vb = engine.topk_rows(b.contiguous(), 10)
        # an index may only differ where the float64 values themselves are closer than the tolerance
        # (near-ties): the float64 value of the node picked at rank p must be within 2 tol of the p-th best
        same = ia == ib
        gap_ok = (vb - torch.gather(b, 1, ia.long())).abs() <= 2 * PARITY_TOL
        top_ok = bool((same | gap_ok).all().item())
        top_same = float(same.double().mean().item())
    del ref
    torch.cuda.empty_cache()
    out = {"max_abs": worst, "rows_checked": rows, "tol": PARITY_TOL, "iterations": iterations,
           "unit_diagonal": diag_ok, "against": "single-GPU float64 CSR path of this engine, same graph, same iterations",
           "topk10_consistent": top_ok, "topk10_identical_frac": top_same if rows else None}
    if world > 1:
        every = [None] * world
        dist.all_gather_object(every, out)
        out = {"max_abs": max(e["max_abs"] for e in every), "rows_checked": sum(e["rows_checked"] for e in every),
               "tol": PARITY_TOL, "iterations": iterations, "unit_diagonal": all(e["unit_diagonal"] for e in every),
               "against": out["against"] + " (every rank checks the rows it owns)",
               "topk10_consistent": all(e["topk10_consistent"] is not False for e in every),
               "per_rank_max_abs": [e["max_abs"] for e in every]}
    if not (out["max_abs"] <= PARITY_TOL and out["unit_diagonal"]):
        raise SystemExit(f"PARITY FAILED: {json.dumps(out)}")
    return out


def run_e2e(args, edges, n, world, rank):
    """``SimRank().fit(DataFrame)`` -> DataFrame with K iterations (eps=0): host graph build, H2D of
    the CSR graph, K device iterations each reading back max|dS|, D2H of the full S."""
    import pandas as pd
    import torch
    from SimRank import SimRank as M
    frm, to = edges
    df = pd.DataFrame({"from": frm, "to": to})
    K = args.steps
    # under torchrun every rank keeps its own row block of the result (gather="local"): the full
    # matrix reaches the host exactly once, spread over the ranks
    obj = M.SimRank(mode=args.mode, slices=args.slices, gather="local" if world > 1 else "all")
    obj.fit(df, iterations=1, eps=0.0, verbose=False)              # warm-up: allocator, pinned pool, library
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    S = obj.fit(df, iterations=K, eps=0.0, verbose=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    assert S.shape[1] == n and obj.fit_info_.applied == K
    rows = torch.tensor([S.shape[0]], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(rows)
    assert int(rows.item()) == n
    h2d = (len(frm) * 4 + (n + 1) * 8 + 2 * n * 8) / K
    d2h = (n * n * 8 / world + 16 * K) / K
    return {"value": K / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "seconds_total": dt, "stages_s": {k: round(v, 4) for k, v in obj.fit_timings_.items()},
            "note": ("whole fit(): pandas graph build + H2D + K iterations + D2H of S into a DataFrame"
                     + ("; every rank returns its own row block (gather='local')" if world > 1 else ""))}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
