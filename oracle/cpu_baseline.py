"""Timing of the reference's CPU iteration (numpy float64, OpenBLAS threads) on a bounded
sample.  TEST/BENCH INFRASTRUCTURE ONLY -- used by bench.py's ``cpu_baseline`` and
``--impl reference`` legs, never by the product path.

The reference cannot execute on this image (see oracle/simrank_oracle.py), so the timed code is
the oracle's restatement of exactly the expressions of SimRank.py:130-140:

    _converged(old_S, new_S, eps)            (abs(s1 - s2) > eps).sum()        SimRank.py:74
    old_S = copy.deepcopy(new_S)                                                SimRank.py:138
    new_S = C * G.dot(new_S).dot(G.T)        two dense dgemm                     SimRank.py:139
    np.fill_diagonal(new_S, 1)                                                  SimRank.py:140

One full iteration at n = 32768 is 1.4e14 flop (~8 min on 8 cores), so a step times a ROW PANEL
of the update -- ``C * G[rows, :].dot(S).dot(G.T)`` for ``rows`` consecutive rows, which is
rows/n of the dgemm work with the same operand shapes and BLAS blocking -- plus the
element-wise passes on the same panel, and scales by n/rows.
"""
from __future__ import annotations

import copy
import os
import time

import numpy as np


def dense_graph(indptr, indices, g, n):
    """Dense float64 G (what the reference keeps in ``self.Graph.values``)."""
    G = np.zeros((n, n), dtype=np.float64)
    deg = np.diff(indptr)
    G[np.repeat(np.arange(n), deg), indices] = np.repeat(g, deg)
    return G


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        th = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        if th:
            return int(max(th))
    except Exception:
        pass
    return os.cpu_count() or 1


def panel_iteration_seconds(G: np.ndarray, S: np.ndarray, rows: int, C: float = 0.8, eps: float = 1e-4) -> float:
    """Wall time of one reference iteration restricted to the first ``rows`` rows of S."""
    n = G.shape[0]
    rows = min(rows, n)
    old = np.zeros((rows, n))
    t0 = time.perf_counter()
    _ = (abs(old - S[:rows]) > eps).sum()                     # SimRank.py:74
    old = copy.deepcopy(S[:rows])                             # SimRank.py:138
    new = C * G[:rows].dot(S).dot(G.T)                        # SimRank.py:139
    new[np.arange(rows), np.arange(rows)] = 1                 # SimRank.py:140 on the panel
    t1 = time.perf_counter()
    assert new.shape == (rows, n) and old.shape == (rows, n)
    return t1 - t0


def iterations_per_second(indptr, indices, g, n, target_seconds=15.0, steps=1, warmup=0, max_bytes=None):
    """-> dict(value=iter/s for the FULL n x n iteration, seconds_per_step, rows, n_timed, note).

    If the two dense n x n float64 operands do not fit in ``max_bytes`` of host memory the
    sample shrinks n (leading principal sub-graph) and extrapolates by (n/n_timed)^3."""
    n_timed = n
    need = 2 * n * n * 8 + (64 << 20)
    note = ""
    if max_bytes is not None and need > max_bytes:
        n_timed = int((max_bytes / 2.5 / 8) ** 0.5) // 1024 * 1024
        n_timed = max(1024, min(n, n_timed))
        note = f"host RAM too small for n={n}: timed the leading {n_timed}-node sub-graph, extrapolated by (n/n_timed)^3; "
    if n_timed != n:
        keep = indices < n_timed
        rows_of = np.repeat(np.arange(n), np.diff(indptr))
        sel = keep & (rows_of < n_timed)
        sub_ptr = np.zeros(n_timed + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows_of[sel], minlength=n_timed)[:n_timed], out=sub_ptr[1:])
        G = dense_graph(sub_ptr, indices[sel], g[:n_timed], n_timed)
    else:
        G = dense_graph(indptr, indices, g, n)
    S = np.full((n_timed, n_timed), 0.01)                     # dgemm time does not depend on the values
    np.fill_diagonal(S, 1.0)
    # calibrate the panel height on a small probe so that one step costs ~target_seconds
    probe_rows = min(n_timed, 64)
    panel_iteration_seconds(G, S, probe_rows)                 # BLAS warm-up
    t_probe = panel_iteration_seconds(G, S, probe_rows)
    rows = int(min(n_timed, max(probe_rows, probe_rows * target_seconds / max(t_probe, 1e-6))))
    rows = max(64, rows // 64 * 64) if n_timed >= 64 else n_timed
    for _ in range(warmup):
        panel_iteration_seconds(G, S, rows)
    times = [panel_iteration_seconds(G, S, rows) for _ in range(max(1, steps))]
    sec_panel = float(np.median(times))
    sec_full_timed_n = sec_panel * n_timed / rows
    sec_full = sec_full_timed_n * (n / n_timed) ** 3
    return dict(value=1.0 / sec_full, seconds_per_step=sec_panel, rows=rows, n_timed=n_timed,
                cores=blas_threads(),
                sample=(f"{note}row panel of {rows}/{n_timed} rows of one iteration "
                        f"(C*G[rows].dot(S).dot(G.T) + converged + deepcopy + fill_diagonal, float64 numpy), "
                        f"{sec_panel:.2f} s per panel, scaled by n/rows"))
