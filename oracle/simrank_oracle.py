"""CPU oracle for the SimRank iteration hot path.  TEST INFRASTRUCTURE ONLY.

This file is a float64 numpy restatement of the algorithms in the reference
``SimRank/SimRank.py`` (abbreviated ``SR.py`` below).  It exists so that the CUDA
path can be checked against the reference's arithmetic; it is NOT part of the
product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.

Why a restatement: the reference's ``fit`` methods cannot execute on this image
(pandas 3: ``DataFrame(index=<set>)`` raises at SR.py:43/188, and the chained
assignment at SR.py:52 is a Copy-on-Write no-op).  The oracle is pinned by
  * the printed outputs of the reference's example notebook (tests/golden/notebook_*.json),
  * outputs of the reference's own ``fit`` loops executed here through a small shim
    (tests/golden/make_golden.py -> tests/golden/ref_*.npz).

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

__all__ = [
    "directed_graph", "bipartite_graph", "evidence", "spread", "weight",
    "converged", "simrank", "simrank_pp", "bipartite_simrank", "bipartite_simrank_pp",
    "apriori_simrank", "bipartite_apriori_simrank", "topk",
    "fit_directed", "fit_bipartite",
]


# --------------------------------------------------------------------------- graph build
def _safe_inverse(x: pd.Series) -> pd.Series:
    """``(1.0 / x).replace([inf, -inf], nan).fillna(0)``  -- SR.py:49, 197-198."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / x.astype("float64")
    return inv.replace([np.inf, -np.inf], np.nan).fillna(0)


def directed_graph(data, weighted=False, from_node_column="from", to_node_column="to",
                   weight_column="weight"):
    """Edge list -> (nodes, G) with ``G[to, from] = 1 / inNeighbors(to)``.

    Follows SR.py:42-52.  ``nodes`` is ``list(set(from.unique()) | set(to.unique()))``
    (SR.py:42: Python set iteration order) and indexes both axes of ``G`` (SR.py:43).
    ``inNeighbors`` is the groupby *sum of weights* (SR.py:45) or *count of from*
    (SR.py:47); the value written for every in-edge of ``to`` is ``1/inNeighbors`` -- the
    numerator is 1, not the edge weight (SR.py:49).  Duplicate (to, from) pairs raise the
    pivot's ValueError (SR.py:50).  The label-aligned row scatter of SR.py:51-52 is done
    here with an index lookup.
    """
    nodes = list(set(data[from_node_column].unique()) | set(data[to_node_column].unique()))
    n = len(nodes)
    if weighted:
        inn = data.groupby(to_node_column)[weight_column].sum().to_frame(name="inNeighbors")
    else:
        inn = data.groupby(to_node_column)[from_node_column].count().to_frame(name="inNeighbors")
    joined = data.join(inn, on=to_node_column)
    joined = joined.assign(_norm=_safe_inverse(joined["inNeighbors"]))
    piv = joined.pivot(index=to_node_column, columns=from_node_column, values="_norm").fillna(0)
    pos = {label: i for i, label in enumerate(nodes)}
    G = np.zeros((n, n), dtype=np.float64)
    rows = np.fromiter((pos[r] for r in piv.index), dtype=np.int64, count=len(piv.index))
    cols = np.fromiter((pos[c] for c in piv.columns), dtype=np.int64, count=len(piv.columns))
    G[np.ix_(rows, cols)] = piv.to_numpy(dtype=np.float64)
    return nodes, G


def bipartite_graph(data, weighted=False, node_group1_column="user", node_group2_column="item",
                    weight_column="weight"):
    """Edge list -> (labels1, labels2, G12, G21).  Follows SR.py:186-200 (== 377-391).

    ``G12[a, b] = 1/deg1(a)`` and ``G21[b, a] = 1/deg2(b)`` where deg is the groupby count
    (SR.py:194-195) or weight sum (SR.py:191-192).  Both matrices come from ``pivot``
    (SR.py:199-200), so rows and columns are in *sorted label order*; ``labels1/2`` are
    those sorted labels (the positional meaning of the result values).  The reference
    labels its outputs with the set-ordered ``NodesGroup1/2`` instead (SR.py:303); that
    relabelling is a host-side policy and not part of the arithmetic.
    """
    g1, g2 = node_group1_column, node_group2_column
    if weighted:
        d1 = data.groupby(g1)[weight_column].sum().to_frame(name="_d1")
        d2 = data.groupby(g2)[weight_column].sum().to_frame(name="_d2")
    else:
        d1 = data.groupby(g1)[g2].count().to_frame(name="_d1")
        d2 = data.groupby(g2)[g1].count().to_frame(name="_d2")
    joined = data.join(d1, on=g1).join(d2, on=g2)
    joined = joined.assign(_w12=_safe_inverse(joined["_d1"]), _w21=_safe_inverse(joined["_d2"]))
    p12 = joined.pivot(index=g1, columns=g2, values="_w12").fillna(0)
    p21 = joined.pivot(index=g2, columns=g1, values="_w21").fillna(0)
    return (list(p12.index), list(p21.index),
            p12.to_numpy(dtype=np.float64), p21.to_numpy(dtype=np.float64))


# --------------------------------------------------------------------------- SimRank++ preprocessing
def evidence(G: np.ndarray) -> np.ndarray:
    """``E = 1 - 0.5 ** ((G>0) @ (G>0).T)``  -- SR.py:315-316 (int64 counts, f64 power)."""
    A = (np.asarray(G) > 0).astype(np.int64)
    # float64 matmul of 0/1 matrices is exact below 2**53 and is BLAS-fast; cast back to int64
    cnt = np.rint(A.astype(np.float64) @ A.astype(np.float64).T).astype(np.int64)
    return 1 - 0.5 ** cnt


def spread(G: np.ndarray) -> np.ndarray:
    """Per-row ``exp(-var)`` with var = sample variance (ddof=1) of the row's nonzeros.

    SR.py:326-332: ``G.replace(0, nan).var(axis=1).fillna(0).apply(exp(-x))``.  pandas'
    ``nanvar`` is two-pass: mean of the non-NaN entries, then the sum of squared
    deviations divided by ``count - 1``; rows with fewer than two nonzeros give NaN -> 0.
    """
    G = np.asarray(G, dtype=np.float64)
    mask = G != 0
    cnt = mask.sum(axis=1)
    tot = np.where(mask, G, 0.0).sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = tot / cnt
        dev = np.where(mask, G - mean[:, None], 0.0)
        var = (dev * dev).sum(axis=1) / (cnt - 1)
    var = np.where(cnt >= 2, var, 0.0)
    var = np.where(np.isnan(var), 0.0, var)
    return np.exp(-var)


def weight(G: np.ndarray) -> np.ndarray:
    """``W = diag(spread) @ G``  -- SR.py:328-333 (done there as a dense n^3 dgemm)."""
    return spread(G)[:, None] * np.asarray(G, dtype=np.float64)


# --------------------------------------------------------------------------- iteration loops
def converged(s1: np.ndarray, s2: np.ndarray, eps: float) -> bool:
    """``(abs(s1 - s2) > eps).sum() == 0``  -- SR.py:74 (NaN compares False => converged)."""
    return not bool((np.abs(s1 - s2) > eps).sum())


def _loop_single(step, n, iterations, eps):
    """Shared skeleton of SR.py:124-140 / 346-362 / 438-454.

    The convergence test runs BEFORE each update (SR.py:130), so ``applied`` updates
    were performed when it returns, and the last pair is never checked when the loop
    runs out (SR.py:129).
    """
    old = np.zeros((n, n))
    new = np.zeros((n, n))
    np.fill_diagonal(new, 1)
    applied, conv = 0, False
    for _ in range(iterations):
        if converged(old, new, eps):
            conv = True
            break
        old = new.copy()                       # SR.py:138 deepcopy
        new = step(new)
        np.fill_diagonal(new, 1)               # SR.py:140
        applied += 1
    return new, applied, conv


def simrank(G, C=0.8, iterations=100, eps=1e-4):
    """``S <- C * G @ S @ G.T; diag <- 1``  -- SR.py:129-140.  Returns (S, applied, converged)."""
    G = np.asarray(G, dtype=np.float64)
    return _loop_single(lambda s: C * G.dot(s).dot(G.T), G.shape[0], iterations, eps)


def simrank_pp(W, E, C=0.8, iterations=100, eps=1e-4):
    """``S <- E * C * W @ S @ W.T; diag <- 1``  -- SR.py:351-362 (evaluation order kept:
    ``(E * C)`` first, then the two products, then the Hadamard)."""
    W = np.asarray(W, dtype=np.float64)
    E = np.asarray(E, dtype=np.float64)
    return _loop_single(lambda s: E * C * W.dot(s).dot(W.T), W.shape[0], iterations, eps)


def apriori_simrank(W, E, prior, C=0.8, lbd=0.5, iterations=100, eps=1e-4):
    """``S <- (1-lbd) * E * C * W S W.T + lbd * prior``  -- SR.py:443-454."""
    W = np.asarray(W, dtype=np.float64)
    E = np.asarray(E, dtype=np.float64)
    prior = np.asarray(prior, dtype=np.float64)
    return _loop_single(lambda s: (1 - lbd) * E * C * W.dot(s).dot(W.T) + lbd * prior,
                        W.shape[0], iterations, eps)


def _loop_pair(step1, step2, n1, n2, iterations, eps):
    """Shared skeleton of SR.py:280-302 / 402-424 / 470-492 (Gauss-Seidel alternation;
    joint convergence test SR.py:289)."""
    old1, new1 = np.zeros((n1, n1)), np.zeros((n1, n1))
    old2, new2 = np.zeros((n2, n2)), np.zeros((n2, n2))
    np.fill_diagonal(new1, 1)
    np.fill_diagonal(new2, 1)
    applied, conv = 0, False
    for _ in range(iterations):
        if converged(old1, new1, eps) and converged(old2, new2, eps):
            conv = True
            break
        old1 = new1.copy()
        new1 = step1(new2)                     # uses the current S2   (SR.py:298)
        np.fill_diagonal(new1, 1)
        old2 = new2.copy()
        new2 = step2(new1)                     # uses the NEW S1       (SR.py:301)
        np.fill_diagonal(new2, 1)
        applied += 1
    return new1, new2, applied, conv


def bipartite_simrank(G12, G21, C1=0.8, C2=0.8, iterations=100, eps=1e-4):
    """SR.py:288-302.  Returns (S1, S2, applied, converged)."""
    G12 = np.asarray(G12, dtype=np.float64)
    G21 = np.asarray(G21, dtype=np.float64)
    return _loop_pair(lambda s2: C1 * G12.dot(s2).dot(G12.T),
                      lambda s1: C2 * G21.dot(s1).dot(G21.T),
                      G12.shape[0], G21.shape[0], iterations, eps)


def pp_group2_evidence(E1, E2):
    """Evidence used for the group-2 update of the bipartite SimRank++ classes.

    SR.py:423 / 491 multiply the group-2 update by ``Evidence_N1``.  With n1 == n2 that
    is what the reference computes and the oracle reproduces it; with n1 != n2 the
    reference raises (numpy broadcast error), so the intended ``Evidence_N2`` is used.
    """
    return E1 if E1.shape == E2.shape else E2


def bipartite_simrank_pp(W1, W2, E1, E2, C1=0.8, C2=0.8, iterations=100, eps=1e-4):
    """SR.py:410-424.  ``S1 <- E1*C1*W1 S2 W1.T``; ``S2 <- E*C2*W2 S1 W2.T`` with
    ``E = pp_group2_evidence(E1, E2)``."""
    W1 = np.asarray(W1, dtype=np.float64)
    W2 = np.asarray(W2, dtype=np.float64)
    Eg2 = pp_group2_evidence(np.asarray(E1), np.asarray(E2))
    return _loop_pair(lambda s2: E1 * C1 * W1.dot(s2).dot(W1.T),
                      lambda s1: Eg2 * C2 * W2.dot(s1).dot(W2.T),
                      W1.shape[0], W2.shape[0], iterations, eps)


def bipartite_apriori_simrank(W1, W2, E1, E2, prior1, prior2, C1=0.8, C2=0.8, lbd1=0.5, lbd2=0.5,
                              iterations=100, eps=1e-4):
    """SR.py:478-492."""
    W1 = np.asarray(W1, dtype=np.float64)
    W2 = np.asarray(W2, dtype=np.float64)
    Eg2 = pp_group2_evidence(np.asarray(E1), np.asarray(E2))
    return _loop_pair(
        lambda s2: (1 - lbd1) * E1 * C1 * W1.dot(s2).dot(W1.T) + lbd1 * np.asarray(prior1),
        lambda s1: (1 - lbd2) * Eg2 * C2 * W2.dot(s1).dot(W2.T) + lbd2 * np.asarray(prior2),
        W1.shape[0], W2.shape[0], iterations, eps)


# --------------------------------------------------------------------------- retrieval
def topk(S: np.ndarray, k: int):
    """Row-wise top-k: ``argsort(-S[i], kind='stable')[:k]`` (ties -> lower index first).

    The reference has no retrieval call (``fit`` returns whole DataFrames, SR.py:141);
    this is the oracle for the engine's top-k helper.
    """
    S = np.asarray(S)
    idx = np.argsort(-S, axis=1, kind="stable")[:, :k]
    return idx, np.take_along_axis(S, idx, axis=1)


# --------------------------------------------------------------------------- whole-fit helpers
def fit_directed(data, kind="simrank", C=0.8, weighted=False, from_node_column="from",
                 to_node_column="to", weight_column="weight", iterations=100, eps=1e-4,
                 prior=None, lbd=0.5):
    """``SimRank().fit`` / ``SimRankPP().fit`` / ``AprioriSimRank().fit`` end to end
    (SR.py:79-141, 339-363, 431-455).  Returns (nodes, S, applied, converged)."""
    nodes, G = directed_graph(data, weighted, from_node_column, to_node_column, weight_column)
    if kind == "simrank":
        S, k, c = simrank(G, C, iterations, eps)
    elif kind == "simrank_pp":
        S, k, c = simrank_pp(weight(G), evidence(G), C, iterations, eps)
    elif kind == "apriori":
        S, k, c = apriori_simrank(weight(G), evidence(G), prior, C, lbd, iterations, eps)
    else:
        raise ValueError(kind)
    return nodes, S, k, c


def fit_bipartite(data, kind="simrank", C1=0.8, C2=0.8, weighted=False, node_group1_column="user",
                  node_group2_column="item", weight_column="weight", iterations=100, eps=1e-4,
                  prior1=None, prior2=None, lbd1=0.5, lbd2=0.5):
    """``BipartiteSimRank().fit`` / ``BipartiteSimRankPP().fit`` / ``BipartitleAprioriSimRank().fit``
    end to end (SR.py:227-303, 393-425, 461-493).  Returns (labels1, labels2, S1, S2, applied, converged)."""
    l1, l2, G12, G21 = bipartite_graph(data, weighted, node_group1_column, node_group2_column,
                                       weight_column)
    if kind == "simrank":
        S1, S2, k, c = bipartite_simrank(G12, G21, C1, C2, iterations, eps)
    elif kind == "simrank_pp":
        S1, S2, k, c = bipartite_simrank_pp(weight(G12), weight(G21), evidence(G12), evidence(G21),
                                            C1, C2, iterations, eps)
    elif kind == "apriori":
        S1, S2, k, c = bipartite_apriori_simrank(weight(G12), weight(G21), evidence(G12), evidence(G21),
                                                 prior1, prior2, C1, C2, lbd1, lbd2, iterations, eps)
    else:
        raise ValueError(kind)
    return l1, l2, S1, S2, k, c
