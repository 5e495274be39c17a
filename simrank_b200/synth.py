"""Seeded synthetic graphs shaped like the five BASELINE.json configs.

There is no network for the BTS-flight / MovieLens files the reference's notebook used, so
every benchmark and parity input is generated here.  Generators return index arrays; the
``*_frame`` helpers wrap them into the edge-list DataFrames the drop-in classes take.
All pairs are unique (the reference's ``pivot`` raises on duplicates, SimRank.py:50).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def _popularity(rng, n, alpha):
    p = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    p /= p.sum()
    return p[rng.permutation(n)]


def _unique_pairs(rng, m, n_rows, n_cols, p_rows, p_cols, forbid_diag):
    """m unique (row, col) pairs, rows ~ p_rows, cols ~ p_cols, drawn independently."""
    have = np.empty(0, dtype=np.int64)
    while have.size < m:
        need = int((m - have.size) * 1.3) + 1024
        r = rng.choice(n_rows, size=need, p=p_rows)
        c = rng.choice(n_cols, size=need, p=p_cols)
        if forbid_diag:
            keep = r != c
            r, c = r[keep], c[keep]
        key = r.astype(np.int64) * n_cols + c
        merged = np.concatenate([have, key])
        _, first = np.unique(merged, return_index=True)
        have = merged[np.sort(first)]            # keep draw order
    have = have[:m]
    return have // n_cols, have % n_cols


def _distinct_per_row(rng, rows, n_cols, k, p):
    """[rows, k] column indices, distinct inside each row, drawn ~ p (vectorised: oversample with
    replacement, keep the first k distinct draws of each row, redraw the few rows that fall short)."""
    out = np.empty((rows, k), dtype=np.int64)
    todo = np.arange(rows)
    width = 2 * k + 8
    while todo.size:
        d = rng.choice(n_cols, size=(todo.size, width), p=p)
        order = np.argsort(d, axis=1, kind="stable")
        sd = np.take_along_axis(d, order, axis=1)
        dup_sorted = np.zeros_like(sd, dtype=bool)
        dup_sorted[:, 1:] = sd[:, 1:] == sd[:, :-1]                # later draws of an already seen value
        dup = np.empty_like(dup_sorted)
        np.put_along_axis(dup, order, dup_sorted, axis=1)
        rank = np.cumsum(~dup, axis=1)                             # 1-based index among the distinct draws
        ok = rank[:, -1] >= k
        take = (~dup) & (rank <= k)
        good = np.nonzero(ok)[0]
        out[todo[good]] = d[good][take[good]].reshape(good.size, k)
        todo = todo[~ok]
        width *= 2
    return out


def directed_edges(n, m, alpha, seed):
    """-> (from_idx, to_idx): m unique directed edges without self-loops."""
    rng = np.random.default_rng(seed)
    p = _popularity(rng, n, alpha)
    uni = np.full(n, 1.0 / n)
    # sources follow the popularity law, targets a flatter one: hubs send to many nodes
    to, frm = _unique_pairs(rng, m, n, n, 0.5 * p + 0.5 * uni, p, forbid_diag=True)
    return frm, to


def bipartite_edges(n1, n2, m, alpha_items, seed, min_per_user=0):
    """-> (user_idx, item_idx): m unique pairs; item popularity ~ rank^-alpha, user activity
    lognormal; every user gets at least ``min_per_user`` items."""
    rng = np.random.default_rng(seed)
    p_items = _popularity(rng, n2, alpha_items)
    act = rng.lognormal(0.0, 1.0, n1)
    p_users = act / act.sum()
    if min_per_user:
        base_u = np.repeat(np.arange(n1), min_per_user)
        base_i = _distinct_per_row(rng, n1, n2, min_per_user, p_items).ravel()
        rest = m - base_u.size
        u, i = _unique_pairs(rng, max(rest, 0) + base_u.size, n1, n2, p_users, p_items, forbid_diag=False)
        key = np.concatenate([base_u.astype(np.int64) * n2 + base_i, u.astype(np.int64) * n2 + i])
        _, first = np.unique(key, return_index=True)
        key = key[np.sort(first)][:m]
        return key // n2, key % n2
    return _unique_pairs(rng, m, n1, n2, p_users, p_items, forbid_diag=False)


# ----------------------------------------------------------------------------- BASELINE configs
CONFIGS = {
    # name: class, shape parameters, run parameters
    "cfg1": dict(cls="SimRank", n=350, m=6000, alpha=1.0, seed=1, C=0.8, iterations=10),
    "cfg2": dict(cls="SimRankPP", n=4096, m=65536, alpha=1.2, seed=2, C=0.8, iterations=10),
    "cfg3": dict(cls="BipartiteSimRank", n1=943, n2=1682, m=100_000, alpha=1.0, seed=3, C1=0.8, C2=0.8,
                 iterations=10),
    "cfg4": dict(cls="SimRank", n=32768, m=2_097_152, alpha=0.5, seed=4, C=0.8, iterations=5),
    "cfg5": dict(cls="BipartiteSimRankPP", n1=138_493, n2=26_744, m=20_000_263, alpha=1.0, seed=5, C1=0.8,
                 C2=0.8, iterations=3),
}


def directed_frame(n, m, alpha, seed, weights=None):
    frm, to = directed_edges(n, m, alpha, seed)
    df = pd.DataFrame({"from": 10000 + frm, "to": 10000 + to})
    rng = np.random.default_rng(seed + 1000)
    if weights == "flights":
        df["weight"] = rng.integers(1, 601, len(df))
    elif weights == "lognormal":
        df["weight"] = rng.lognormal(0.0, 1.0, len(df))
    return df


def bipartite_frame(n1, n2, m, alpha, seed, ratings="stars", min_per_user=0):
    u, i = bipartite_edges(n1, n2, m, alpha, seed, min_per_user)
    df = pd.DataFrame({"user": 1 + u, "item": 1 + i})
    rng = np.random.default_rng(seed + 1000)
    if ratings == "stars":
        df["weight"] = rng.choice([1, 2, 3, 4, 5], len(df), p=[0.06, 0.11, 0.27, 0.34, 0.22]).astype(np.float64)
    elif ratings == "half_stars":
        df["weight"] = rng.choice(np.arange(0.5, 5.01, 0.5), len(df))
    return df


def config_frame(name, scale=1.0):
    """Edge-list DataFrame for BASELINE config ``name`` (optionally shrunk by ``scale`` in every
    dimension, edges scaled to keep the mean degree)."""
    c = CONFIGS[name]
    if "n" in c:
        n = max(8, int(c["n"] * scale))
        m = max(16, int(c["m"] * scale))
        w = {"cfg1": "flights", "cfg2": "lognormal"}.get(name)
        return directed_frame(n, m, c["alpha"], c["seed"], w)
    n1, n2 = max(8, int(c["n1"] * scale)), max(8, int(c["n2"] * scale))
    m = max(32, int(c["m"] * scale))
    r = "stars" if name == "cfg3" else "half_stars"
    return bipartite_frame(n1, n2, m, c["alpha"], c["seed"], r, min_per_user=20 if scale == 1.0 else 0)
