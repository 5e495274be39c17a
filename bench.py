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
    ap.add_argument("--mode", default=None, choices=["i8", "csr", "csr16"],
                    help="dense tensor-core chain (graded), the float64 CSR SpMM path, or the fixed-point CSR SpMM path")
    ap.add_argument("--slices", default="auto", help="uint8 planes per matrix: 2, 3, 4 or auto (error-bound driven)")
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg5"],
                    help="cfg4: the headline (directed, n = 32768); cfg5: BipartitleSimRankPP on the MovieLens-20M-shaped "
                         "graph, row-sharded, CSR SpMM path (BASELINE configs[4]; run under torchrun on 8 GPUs)")
    ap.add_argument("--scale", type=float, default=1.0, help="cfg5 only: shrink every dimension (smoke runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "i8", "csr", "csr16"],
                    help="mode= of the drop-in class in the end-to-end measurement; 'auto' is what a user's "
                         "SimRank().fit(df) runs (engine.choose_mode picks the fastest path that keeps 1e-6)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-csr", action="store_true", help="skip the secondary CSR SpMM measurement of the default run")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the float64 re-run that checks the timed solver's result (outside the timed region)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    a = ap.parse_args()
    if a.mode is None:
        a.mode = "i8" if a.config == "cfg4" else "csr16"
    return a


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
def int8_peak():
    """Measured tcgen05 kind::i8 peak of this pool's B200s (scripts/micro/i8_peak.cu, operands resident in
    shared memory): profiles/int8_peak.json, else None."""
    path = os.path.join(ROOT, "profiles", "int8_peak.json")
    return json.load(open(path)) if os.path.exists(path) else None


def measure(args, op, dev, mode, world, local):
    """W warm-up + K timed iterations of one solver family; -> the JSON fragment of that path and the
    finished solver (for the parity check)."""
    import torch
    import torch.distributed as dist
    from simrank_b200 import _lib as _srk_lib
    from simrank_b200 import engine
    n = op.M
    if world > 1:
        from simrank_b200 import dist as sdist
        solver = sdist.ShardedDirectedSolver(op, 0.8, mode=mode, ns=args.slices, device=dev)
        halves = solver.halves
    else:
        solver = engine.DirectedSolver(engine.DeviceOperator(op, dev), 0.8, mode=mode, ns=args.slices)
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
    launches_before = _srk_lib.LAUNCHES
    with ClockSampler(local) as clocks:
        t0.record()
        last = None
        for _ in range(args.steps):
            last = solver.step()
        t1.record()
        sync_all()
    launches = _srk_lib.LAUNCHES - launches_before          # library calls (>= 1 kernel each) enqueued by the timed steps
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
    if mode == "i8":
        gemms = {k: v for k, v in kernels.items() if "half" in k}
        dom = max(gemms, key=lambda k: gemms[k]["ms"])
        ach = flops_half / (kernels[dom]["ms"] * 1e-3) / 1e12
        sym = world == 1 and dom.endswith("final")
        executed = ach * used[-1] * (0.5 if sym else 1.0)
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["tensor_sustained"], "traffic": ncu_traffic(dom, used[-1], n, world),
                "peak_kind": f"dense bf16 sustained, {pk['source']}; burst {pk['tensor_burst']}",
                "executed_int8_tops": executed,
                "tensor_pipe_active_ncu": ncu_pipe_active(dom, used[-1], n, world),
                "note": ("achieved = algorithmic 2n^3 flop of one half-product / mean launch time inside the timed "
                         f"steps; the kernel executes {used} u8 x u8 -> s32 tcgen05 products per algorithmic one"
                         + ("; this launch computes only the upper triangle of the symmetric result" if sym else ""))}
        i8 = int8_peak()
        if i8:
            # the pipe this kernel actually runs on: executed u8 x u8 -> s32 MACs against the measured kind::i8 peak
            roof["frac_int8"] = executed / i8["int8_tops"]
            roof["peak_int8"] = {"value": i8["int8_tops"], "unit": "Top/s", "sm_mhz": i8["sm_mhz"], "source": i8["source"]}
    else:
        halves_only = {k: v for k, v in kernels.items() if "half" in k}
        dom = max(halves_only, key=lambda k: halves_only[k]["ms"])
        by = (3 if dom.endswith("final") else 2) * n * n * 8.0 / world
        ach = by / (kernels[dom]["ms"] * 1e-3) / 1e9
        it_bytes = 5.0 * n * n * 8.0 / world
        nnz = float(op.nnz)
        esz = 2.0 if mode == "csr16" else 8.0
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": ncu_traffic(dom, 0 if mode == "csr" else 16, n, world),
                "peak_kind": pk["source"],
                "note": "algorithmic bytes (SURVEY.md 8d): first half 2 n^2 s, second half 3 n^2 s (s = 8)",
                "iteration": {"algorithmic_bytes": it_bytes, "achieved": it_bytes / (ms * 1e-3) / 1e9,
                              "frac": it_bytes / (ms * 1e-3) / 1e9 / pk["hbm"]},
                # what actually bounds the half-products: rows of X gathered through L2 (nnz * n elements
                # per half; the symmetric second half gathers about half of that)
                "l2_gather": {"bytes_first_half": nnz * n * esz / world,
                              "achieved_first_half_GBs": (nnz * n * esz / world) /
                              (kernels[[k for k in halves_only if k.endswith("first")][0]]["ms"] * 1e-3) / 1e9,
                              "roof_GBs": 15500.0, "roof_source": "profiles/r2_micro_tma_gather_rate.txt (ldg, 1 KB segments)"}}

    frag = {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms,
            "dtype": {"i8": f"u8 x u8 -> s32 ({used} fixed-point planes), f64 epilogue", "csr": "f64",
                      "csr16": "u16 gather -> exact u32 sums, f64 epilogue"}[mode],
            "mode": mode, "slices_used": used,
            "algorithmic_tflops": 4.0 * n ** 3 / (ms * 1e-3) / 1e12,
            "roofline": roof, "kernels": kernels, "gpu_launches": launches,
            "last_maxdiff": last if not isinstance(last, tuple) else list(last)}
    if not args.no_parity:
        frag["parity"] = run_parity(args, solver, halves, op, dev, world, args.warmup + args.steps)
    clk = clocks.summary()
    if world > 1:
        # the per-kernel times above are rank 0's; the spread over the ranks shows who waits for whom
        every = [None] * world
        dist.all_gather_object(every, {k: v["ms"] for k, v in kernels.items()})
        frag["kernels_rank_spread_ms"] = {k: [round(min(e.get(k, 0.0) for e in every), 3),
                                              round(max(e.get(k, 0.0) for e in every), 3)] for k in kernels}
        gathered = [None] * world
        dist.all_gather_object(gathered, clk)
        mhz = [g["sm_mhz"] for g in gathered if g["sm_mhz"]]
        clk = {"sm_mhz": min(mhz) if mhz else None, "sm_max_mhz": clk["sm_max_mhz"],
               "reasons": sorted(set(r for g in gathered for r in g["reasons"]))}
    frag["clocks"] = clk
    del solver, halves
    torch.cuda.empty_cache()
    return frag


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

    frag = measure(args, op, dev, args.mode, world, local)
    line = {"metric": METRIC, "value": frag["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": frag["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": frag["dtype"], "data": "synthetic",
            "config": {"workload": name, "mode": args.mode, "slices": args.slices or "auto",
                       "slices_used": frag["slices_used"],
                       "l2": "operands (>= 1 GB) are far larger than the 126 MB L2; no flush needed",
                       "parallelism": f"S row-sharded over {world} GPU(s)"}}
    for k in ("algorithmic_tflops", "roofline", "kernels", "gpu_launches", "last_maxdiff", "parity",
              "kernels_rank_spread_ms", "clocks"):
        if k in frag:
            line[k] = frag[k]

    # ---- the CSR SpMM path on the same workload (BASELINE cfg4: "also report CSR path"): fixed-point
    # gather (csr16) when the graded path is the dense chain
    if args.mode == "i8" and not args.no_csr:
        other = measure(args, op, dev, "csr16", world, local)
        line["csr_path"] = {k: other[k] for k in ("mode", "value", "unit", "ms_per_step", "dtype", "roofline", "kernels",
                                                  "gpu_launches", "parity", "clocks") if k in other}

    # ---- end to end through the public API: DataFrame in -> DataFrame out, host buffers
    if not args.no_e2e:
        line["e2e"] = run_e2e(args, edges, n, world, rank)
        torch.cuda.empty_cache()
        if line["e2e"]["mode_used"] != args.mode:          # the same call pinned to the path `value` is measured on
            pinned = run_e2e(args, edges, n, world, rank, mode=args.mode)
            line["e2e"]["with_mode_of_value"] = {k: pinned[k] for k in ("value", "mode_used", "seconds_total", "stages_s")}
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


def run_e2e(args, edges, n, world, rank, mode=None):
    """``SimRank().fit(DataFrame)`` -> DataFrame with K iterations (eps=0): host graph build, H2D of
    the CSR graph, K device iterations each reading back max|dS|, D2H of the full S.  ``mode`` is the
    constructor's (default: the class default 'auto', the call a user makes)."""
    import pandas as pd
    import torch
    from SimRank import SimRank as M
    frm, to = edges
    df = pd.DataFrame({"from": frm, "to": to})
    K = args.steps
    # under torchrun every rank keeps its own row block of the result (gather="local"): the full
    # matrix reaches the host exactly once, spread over the ranks
    mode = mode or args.e2e_mode
    obj = M.SimRank(mode=mode, slices=args.slices, gather="local" if world > 1 else "all")
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
            "mode": mode, "mode_used": obj.fit_info_.mode, "seconds_total": dt, "stages_s": {k: round(v, 4) for k, v in obj.fit_timings_.items()},
            "note": ("whole fit(): pandas graph build + H2D + K iterations + D2H of S into a DataFrame"
                     + ("; every rank returns its own row block (gather='local')" if world > 1 else ""))}


# ----------------------------------------------------------------------------- BASELINE configs[4]
def run_cfg5(args):
    """BipartitleSimRankPP on the synthetic MovieLens-20M-shaped graph (138 493 x 26 744, 20 000 263 weighted
    ratings), S1 / S2 row-sharded over the GPUs of the box, CSR SpMM path (fixed-point gather by default).
    A step = one Gauss-Seidel iteration (S1 update, then S2 update from the new S1; SimRank.py:410-424) with
    both max|dS| read-backs.  Parity: the same iterations re-run in float64 on the same sharded solver
    (every update forced to the float64 gather), ALL local rows of both matrices compared."""
    import torch
    import torch.distributed as dist
    from simrank_b200 import drivers, engine, synth
    args.slices = None if args.slices in ("auto", None) else int(args.slices)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = engine.require_cuda(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.perf_counter()
    df = synth.config_frame("cfg5", args.scale)
    t_data = time.perf_counter() - t0

    def build(mode):
        _, _, l1, l2, op12, op21 = drivers.build_bipartite(df, True, "user", "item", "weight")
        W1, W2 = drivers.weight(op12), drivers.weight(op21)
        E1, E2 = drivers.evidence(op12), drivers.evidence(op21)
        return drivers.bipartite_solver(W1, W2, 0.8, 0.8, evidence1=E1, evidence2=E2, mode=mode, slices=args.slices), op12

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    t0 = time.perf_counter()
    solver, op12 = build(args.mode)
    sync_all()
    t_setup = time.perf_counter() - t0
    n1, n2, nnz = op12.M, op12.K, op12.nnz
    halves = [solver.h1, solver.h2]
    for _ in range(args.warmup):
        solver.step()
    for h in halves:
        h.events = []
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from simrank_b200 import _lib as _srk_lib
    launches_before = _srk_lib.LAUNCHES
    wall0 = time.perf_counter()
    with ClockSampler(local) as clocks:
        e0.record()
        last = None
        for _ in range(args.steps):
            last = solver.step()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - wall0) / args.steps      # every step ends with its scalar read-back
        sync_all()
    launches = _srk_lib.LAUNCHES - launches_before
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    kernels = {}
    for name, h in (("S1", solver.h1), ("S2", solver.h2)):
        per = {}
        for nm, a, b in h.events:
            per.setdefault(nm, []).append(a.elapsed_time(b))
        kernels[name] = {k: {"ms": statistics.mean(v), "launches": len(v)} for k, v in per.items()}
        h.events = None
    used = {name: sorted(set(getattr(h, "slices_used", [])[-args.steps:])) for name, h in (("S1", solver.h1), ("S2", solver.h2))}
    pk = peaks()
    # SURVEY.md 8d, bipartite CSR SpMM: s (3 n1^2 + 3 n2^2 + 4 n1 n2) + e (n1^2 + n2^2), s = 8, e = 1
    alg_bytes = 8.0 * (3.0 * n1 * n1 + 3.0 * n2 * n2 + 4.0 * n1 * n2) + 1.0 * (n1 * n1 + n2 * n2)
    esz = {"csr16": 2.0, "csr": 8.0}.get(args.mode)
    gather = 2.0 * nnz * (n1 + n2) * esz if esz else None
    kernel_ms = sum(v["ms"] for d in kernels.values() for k, v in d.items() if "half" in k)
    roof = {"bound": "hbm", "kernel": "iteration (4 half-products + quantisers)", "achieved": alg_bytes / (ms * 1e-3) / 1e9,
            "peak": pk["hbm"] * world, "unit": "GB/s", "frac": alg_bytes / (ms * 1e-3) / 1e9 / (pk["hbm"] * world),
            "traffic": None, "peak_kind": f"{world} x {pk['source']}",
            "note": "algorithmic bytes per iteration (SURVEY.md 8d bipartite CSR row): s (3 n1^2 + 3 n2^2 + 4 n1 n2) + e (n1^2 + n2^2)"}
    if gather:
        roof["l2_gather"] = {"bytes_per_iteration": gather, "achieved_GBs_per_gpu": gather / world / (kernel_ms * 1e-3) / 1e9,
                             "roof_GBs_per_gpu": 15500.0,
                             "note": "what bounds the half-products: rows gathered through L2, 2 nnz (n1 + n2) elements per iteration"}
    line = {"metric": "simrank_iterations_per_sec_cfg5", "value": 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"i8": "u8 x u8 -> s32 fixed-point planes, f64 epilogue", "csr": "f64",
                      "csr16": "u16 gather -> exact u32 sums, f64 epilogue"}[args.mode],
            "data": "synthetic",
            "config": {"workload": (f"cfg5: BipartitleSimRankPP on a synthetic MovieLens-20M-shaped bipartite graph, {n1} x {n2}, "
                                    f"{nnz} weighted ratings (scale {args.scale}), C1 = C2 = 0.8, Evidence_N2 for group 2"),
                       "mode": solver.mode, "slices_used": used, "parallelism": f"S1 and S2 row-sharded over {world} GPU(s)",
                       "l2": "operands are far larger than the 126 MB L2; no flush needed"},
            "algorithmic_tflops_dense_equivalent": 4.0 * n1 * n2 * (n1 + n2) / (ms * 1e-3) / 1e12,
            "roofline": roof, "kernels": kernels, "gpu_launches": launches,
            "host_overhead_ms_per_step": wall * 1e3 - ms, "seconds": {"data": t_data, "graph_and_setup": t_setup},
            "last_maxdiff": list(last)}
    clk = clocks.summary()
    mem = torch.tensor([torch.cuda.max_memory_allocated() / 2 ** 30], device=dev)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, clk)
        mhz = [g["sm_mhz"] for g in gathered if g["sm_mhz"]]
        clk = {"sm_mhz": min(mhz) if mhz else None, "sm_max_mhz": clk["sm_max_mhz"],
               "reasons": sorted(set(r for g in gathered for r in g["reasons"]))}
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    line["clocks"] = clk
    line["max_gpu_mem_gib"] = round(float(mem.item()), 1)

    # ---- retrieval: top-10 of the local rows to the host (S1 is 153 GB: the matrix itself stays on the GPUs)
    t0 = time.perf_counter()
    idx, vals = engine.topk_rows(solver.h1.local_result().contiguous(), 10)
    idx_h, vals_h = idx.cpu(), vals.cpu()
    line["topk10_seconds"] = time.perf_counter() - t0

    if not args.no_parity:
        mine = [h.local_result().clone() for h in halves]
        total = args.warmup + args.steps
        del solver, halves
        torch.cuda.empty_cache()
        ref, _ = build("csr16" if args.mode != "csr" else "csr")
        for h in (ref.h1, ref.h2):
            h.force_f64 = True                           # every update in float64 (exact sums), same sharding
        for _ in range(total):
            ref.step()
        worst, diag_ok = 0.0, True
        for got, h in zip(mine, (ref.h1, ref.h2)):
            want = h.local_result()
            if got.numel():
                worst = max(worst, float((got - want).abs().max().item()))
                r = torch.arange(h.rows, device=dev)
                diag_ok = diag_ok and bool((got[r, h.row0 + r] == 1.0).all().item())
        out = {"max_abs": worst, "rows_checked": int(sum(m.shape[0] for m in mine)), "tol": PARITY_TOL, "iterations": total,
               "unit_diagonal": diag_ok,
               "against": "the same row-sharded solver with every update forced to the float64 CSR gather; all local rows of S1 and S2"}
        if world > 1:
            every = [None] * world
            dist.all_gather_object(every, out)
            out.update(max_abs=max(e["max_abs"] for e in every), rows_checked=sum(e["rows_checked"] for e in every),
                       unit_diagonal=all(e["unit_diagonal"] for e in every))
        line["parity"] = out
        if not (out["max_abs"] <= PARITY_TOL and out["unit_diagonal"]):
            raise SystemExit(f"PARITY FAILED: {json.dumps(out)}")
    line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": "n/a: the reference raises at SimRank.py:423 for n1 != n2 and S1 alone is 153 GB in float64"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == "cfg5":
        run_cfg5(a)
    else:
        run_engine(a)
