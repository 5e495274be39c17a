#!/usr/bin/env python
"""One rank's CSR half-product launches of a row-sharded fit, replayed on ONE GPU with random data:
lets the kernels be tuned at the shapes of an 8-GPU run (BASELINE cfg5: 138 493 x 26 744, 20M ratings;
cfg4 on 8 ranks) without paying for eight GPUs.  Prints one JSON line per case.

    python scripts/csr_shape_bench.py cfg5_s1_final cfg5_s2_first cfg4_n8_final cfg4_n8_first
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simrank_b200 import _lib, engine, graph  # noqa: E402


def random_bipartite(n1, n2, m, seed):
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, n2 + 1)                               # item popularity ~ 1/rank, as cfg5
    item = rng.choice(n2, size=int(m * 1.15), p=pop / pop.sum())
    user = rng.integers(0, n1, size=item.size)
    key = np.unique(user.astype(np.int64) * n2 + item)[:m]
    return key // n2, key % n2


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def split_info(sp):
    if sp is None:
        return None
    return {"rows": sp.rows, "pieces": sp.pieces, "ranges": sp.ranges, "piece": sp.piece, "min_deg": sp.min_deg}


def args_for(dop, elem, mode):
    a = _lib.CsrArgs()
    a.elem, a.mode = elem, mode
    a.indptr, a.indices, a.g = dop.indptr.data_ptr(), dop.indices.data_ptr(), dop.g.data_ptr()
    a.M = dop.M
    return a


def final_case(dev, op, rows_local, name):
    """Second half on the received panel T[:, rows_out_p]: X [n_in x rows_local] uint16, all graph rows."""
    dop = engine.DeviceOperator(op, dev)
    n_out, n_in = op.M, op.K
    per = engine._round_up(rows_local, 16)
    ld = engine._round_up(n_out, 16)
    X = torch.randint(0, 30000, (n_in, per), dtype=torch.int16, device=dev)
    S = torch.rand((per, ld), dtype=torch.float64, device=dev)
    cnt = torch.randint(0, 30, (per, ld), dtype=torch.int16, device=dev)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    deg = torch.from_numpy(op.deg.astype(np.float64)).to(dev)
    b = args_for(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FINAL)
    b.row_begin, b.row_end = 0, n_out
    b.X, b.ldx, b.L, b.K = X.data_ptr(), per, rows_local, n_in
    b.OUT, b.ldo = S.data_ptr(), ld
    b.in_unit = _lib.RowBound.of(deg.data_ptr(), 1e-9, 0.0)
    b.g_col = dop.g.data_ptr()
    b.counts, b.ld_counts, b.counts_bits, b.add_counts, b.use_evidence = cnt.data_ptr(), ld, 16, 1, 1
    b.epi.coef = 0.8
    b.epi.s_old, b.epi.ld_s_old = S.data_ptr(), ld
    b.epi.maxdiff, b.epi.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    lib = _lib.load()
    via_accum = os.environ.get("SRK_FINAL_VIA_ACCUM", "1") == "1"
    split = engine.ListSplit.plan(dop.indptr, dop.indices, n_in, int(op.deg.max()), all_rows=via_accum)

    def run():
        if split is not None:
            split.accumulate(lib, b.indices, b.X, b.ldx, b.L, b.K, 65535.0)
            if via_accum:
                b.mode, b.accum, b.ld_accum = _lib.SRK_CSR_FINISH, split._accum.data_ptr(), split._accum.shape[1]
            else:
                split.attach(b)
        _lib.check(lib.srk_csr_half(C.byref(b), engine._stream()))
    ms = timed(run)
    ms_accum = timed(lambda: split.accumulate(lib, b.indices, b.X, b.ldx, b.L, b.K, 65535.0)) if split is not None else 0.0
    gather = op.nnz * rows_local * 2.0
    return {"case": name, "ms": ms, "ms_accum": ms_accum, "split": split_info(split), "gather_GB": gather / 1e9, "gather_TBs": gather / ms / 1e9,
            "panel_MB_per_1KB_segment": n_in * 1024 / 1e6, "shape": [n_out, n_in, rows_local]}


def first_case(dev, op, rows_src, world, name):
    """First half: one launch per destination rank over its block of graph rows, X = the transposed local
    rows of S_in [n_in x rows_src] uint16."""
    dop = engine.DeviceOperator(op, dev)
    n_out, n_in = op.M, op.K
    per_out = engine._round_up(-(-n_out // world), 16)
    ldxt = engine._round_up(rows_src, 64)
    X = torch.randint(0, 30000, (n_in, ldxt), dtype=torch.int16, device=dev)
    unit = torch.rand(ldxt, dtype=torch.float64, device=dev) * 1e-9
    deg = torch.from_numpy(op.deg.astype(np.float64)).to(dev)
    send = torch.zeros((world, engine._round_up(rows_src, 16), per_out), dtype=torch.int16, device=dev)
    lib = _lib.load()
    via_accum = os.environ.get("SRK_FIRST_VIA_ACCUM", "1") == "1" and n_in * 1024 <= 64 * 2 ** 20
    split = engine.ListSplit.plan(dop.indptr, dop.indices, n_in, int(op.deg.max()), all_rows=via_accum)

    def run():
        if split is not None:
            split.accumulate(lib, dop.indices.data_ptr(), X.data_ptr(), ldxt, rows_src, n_in, 65535.0)
        for p in range(world):
            lo, hi = min(n_out, p * per_out), min(n_out, (p + 1) * per_out)
            if hi <= lo:
                continue
            a = args_for(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FIRST)
            a.row_begin, a.row_end = lo, hi
            a.X, a.ldx, a.L, a.K = X.data_ptr(), ldxt, rows_src, n_in
            a.OUT, a.ldo = send[p].data_ptr() - 2 * lo, per_out
            a.in_unit = _lib.RowBound.of(unit.data_ptr(), 1.0, 0.0)
            a.out_bound = _lib.RowBound.of(deg.data_ptr(), 1e-4, 0.0)
            if split is not None and via_accum:
                a.mode, a.accum, a.ld_accum = _lib.SRK_CSR_FINISH_FIRST, split._accum.data_ptr(), split._accum.shape[1]
            elif split is not None:
                split.attach(a)
            _lib.check(lib.srk_csr_half(C.byref(a), engine._stream()))
    ms = timed(run)
    ms_accum = timed(lambda: split.accumulate(lib, dop.indices.data_ptr(), X.data_ptr(), ldxt, rows_src, n_in, 65535.0)) \
        if split is not None else 0.0
    gather = op.nnz * rows_src * 2.0
    return {"case": name, "ms": ms, "ms_accum": ms_accum, "split": split_info(split), "gather_GB": gather / 1e9, "gather_TBs": gather / ms / 1e9,
            "panel_MB_per_1KB_segment": n_in * 1024 / 1e6, "shape": [n_out, n_in, rows_src]}


def main():
    dev = engine.require_cuda()
    cases = sys.argv[1:] or ["cfg5_s1_final", "cfg5_s2_first", "cfg4_n8_final", "cfg4_n8_first"]
    ops = {}
    # SRK_SWEEP="A=1,B=2;A=3": every case runs once per ';'-separated environment setting (one graph build)
    sweep = [dict(kv.split("=") for kv in part.split(",") if kv) for part in os.environ.get("SRK_SWEEP", "").split(";")]
    for env in sweep:
        os.environ.update(env)
        for c in cases:
            if c.startswith("cfg5") and "cfg5" not in ops:
                if os.environ.get("SRK_REAL_CFG5", "0") == "1":        # the bench's own synthetic graph (36 s to build)
                    from simrank_b200 import synth
                    c5 = synth.CONFIGS["cfg5"]
                    u, i = synth.bipartite_edges(c5["n1"], c5["n2"], c5["m"], c5["alpha"], c5["seed"], 20)
                else:
                    u, i = random_bipartite(138493, 26744, 20000263, 5)
                ops["cfg5"] = (graph.operator_from_edges(u, i, 138493, 26744), graph.operator_from_edges(i, u, 26744, 138493))
            if c.startswith("cfg4") and "cfg4" not in ops:
                from simrank_b200 import synth
                frm, to = synth.directed_edges(32768, 32768 * 64, 0.5, 4)
                ops["cfg4"] = graph.operator_from_edges(to, frm, 32768, 32768)
            if c == "cfg5_s1_final":
                out = final_case(dev, ops["cfg5"][0], 17312, c)
            elif c == "cfg5_s2_final":
                out = final_case(dev, ops["cfg5"][1], 3344, c)
            elif c == "cfg5_s2_first":
                out = first_case(dev, ops["cfg5"][1], 17312, 8, c)
            elif c == "cfg5_s1_first":
                out = first_case(dev, ops["cfg5"][0], 3344, 8, c)
            elif c == "cfg4_n8_final":
                out = final_case(dev, ops["cfg4"], 4096, c)
            elif c == "cfg4_n8_first":
                out = first_case(dev, ops["cfg4"], 4096, 8, c)
            else:
                raise SystemExit(f"unknown case {c}")
            out["flags"] = os.environ.get("SRK_CSR_FLAGS", "")
            out["env"] = env
            print(json.dumps(out), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
