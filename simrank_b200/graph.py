"""Host-side graph construction: edge-list DataFrame -> row-scaled 0/1 operator in CSR form.

This is the host half of the reference's ``_create_graph`` methods (SimRank.py:24-52,
168-200, 376-391) restated so that it runs on pandas 3 and emits ``G = diag(g) * A`` as
(indptr, indices, g) instead of a dense n x n DataFrame.  The reference can only produce
matrices of that form: the value written for every in-edge of a node is ``1/inNeighbors``
(SimRank.py:49, 197-198).  Semantics kept:

* node order of the directed classes = iteration order of
  ``set(from.unique()) | set(to.unique())`` evaluated here with the same expression
  (SimRank.py:42) -- it is interpreter-defined, so it is never re-derived elsewhere;
* bipartite matrices are in sorted-label order (they come from ``pivot``, SimRank.py:199-200);
* ``inNeighbors`` = groupby sum of weights / count of the partner column (NaN skipped, as
  pandas does), ``1/x`` with +-inf -> 0 (SimRank.py:45-49);
* duplicate (row, column) pairs raise the ``pivot`` ValueError (SimRank.py:50).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import pandas as pd

_DUPLICATE_MSG = "Index contains duplicate entries, cannot reshape"


@dataclass
class HostOperator:
    """``G = diag(g) * A`` with A in CSR (column indices sorted inside each row)."""
    M: int
    K: int
    indptr: np.ndarray          # int64 [M+1]
    indices: np.ndarray         # int32 [nnz]
    g: np.ndarray               # float64 [M]
    deg: np.ndarray = field(default=None)   # int64 [M] structural row degree
    dev_csr: object = field(default=None, repr=False, compare=False)   # (indptr, indices) device tensors, if built there

    def __post_init__(self):
        if self.deg is None:
            self.deg = np.diff(self.indptr).astype(np.int64)

    @property
    def nnz(self) -> int:
        return int(self.indptr[-1])

    @property
    def dead(self) -> np.ndarray:
        """Rows for which ``G > 0`` is False everywhere (SimRank.py:315)."""
        return (~(self.g > 0)).astype(np.uint8)

    def scaled(self, factor: np.ndarray) -> "HostOperator":
        """``diag(factor) * G`` -- the SimRank++ weight matrix W (SimRank.py:333)."""
        return HostOperator(self.M, self.K, self.indptr, self.indices, self.g * factor, self.deg)

    def to_dense(self) -> np.ndarray:
        out = np.zeros((self.M, self.K), dtype=np.float64)
        rows = np.repeat(np.arange(self.M), self.deg)
        out[rows, self.indices] = np.repeat(self.g, self.deg)
        return out


def _inverse_or_zero(x: np.ndarray) -> np.ndarray:
    """``(1.0 / x).replace([inf, -inf], nan).fillna(0)`` (SimRank.py:49)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / np.asarray(x, dtype=np.float64)
    inv[~np.isfinite(inv)] = 0.0
    return inv


def _csr(rows: np.ndarray, cols: np.ndarray, M: int, K: int):
    """Host CSR build (small graphs, and machines without the device builder): one sort of packed keys."""
    if rows.size:
        # one sort of packed (row, column) keys; both halves come back with shifts.  32-bit keys when
        # they fit (n < 65536 on both sides): the sort is the dominant cost and 1.4x faster on uint32
        kb = max(1, int(K - 1).bit_length())
        small = int(max(M - 1, 0)).bit_length() + kb <= 32
        kt = np.uint32 if small else np.int64
        key = (rows.astype(kt) << kt(kb)) | cols.astype(kt)
        key.sort()
        if key.size > 1 and np.any(key[1:] == key[:-1]):
            raise ValueError(_DUPLICATE_MSG)
        rows = key >> kt(kb)
        cols = key & kt((1 << kb) - 1)
    indptr = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=indptr[1:])
    return indptr, cols.astype(np.int32)


def _per_node(series: pd.Series, labels: pd.Index) -> np.ndarray:
    """Align a groupby result (indexed by label) to ``labels``; missing -> NaN."""
    return series.reindex(labels).to_numpy(dtype=np.float64)


def _build_csr(rows, cols, M, K, csr):
    """(indptr, indices, device arrays or None) through ``csr`` (a device builder such as
    engine.device_csr: the GPU replacement of pivot + row scatter) or the host sort."""
    if csr is not None:
        out = csr(rows, cols, M, K)
        if out is not None:
            return out
    indptr, indices = _csr(rows, cols, M, K)
    return indptr, indices, None


def build_directed(data: pd.DataFrame, weighted: bool, from_node_column: str, to_node_column: str,
                   weight_column: str, csr=None):
    """-> (node_set, node_list, HostOperator) for SimRank / SimRankPP / AprioriSimRank.

    ``G[to, from] = 1 / inNeighbors(to)`` (SimRank.py:42-52)."""
    node_set = set(data[from_node_column].unique()) | set(data[to_node_column].unique())
    nodes = list(node_set)
    n = len(nodes)
    labels = pd.Index(nodes)
    rows = labels.get_indexer(data[to_node_column])
    cols = labels.get_indexer(data[from_node_column])
    if weighted:
        inn = _per_node(data.groupby(to_node_column)[weight_column].sum(), labels)
    elif data[from_node_column].isna().any() or data[to_node_column].isna().any():
        inn = _per_node(data.groupby(to_node_column)[from_node_column].count(), labels)
    else:
        # `groupby(to)[from].count()` without missing values is the number of rows per `to` label;
        # nodes that never occur as `to` have no group: NaN after the join (SimRank.py:48), g = 0
        inn = np.bincount(rows, minlength=n).astype(np.float64)
        inn[inn == 0] = np.nan
    g = _inverse_or_zero(inn)
    indptr, indices, dev = _build_csr(rows, cols, n, n, csr)
    op = HostOperator(n, n, indptr, indices, g)
    op.dev_csr = dev
    return node_set, nodes, op


def build_bipartite(data: pd.DataFrame, weighted: bool, node_group1_column: str,
                    node_group2_column: str, weight_column: str, csr=None):
    """-> (set1, set2, sorted1, sorted2, G12, G21) for the bipartite classes.

    ``G12[a, b] = 1/deg1(a)``, ``G21[b, a] = 1/deg2(b)`` in sorted-label order
    (SimRank.py:186-200)."""
    c1, c2 = node_group1_column, node_group2_column
    set1, set2 = set(data[c1].unique()), set(data[c2].unique())
    l1 = pd.Index(data[c1].unique()).sort_values()
    l2 = pd.Index(data[c2].unique()).sort_values()
    if weighted:
        d1 = data.groupby(c1)[weight_column].sum()
        d2 = data.groupby(c2)[weight_column].sum()
    else:
        d1 = data.groupby(c1)[c2].count()
        d2 = data.groupby(c2)[c1].count()
    g1 = _inverse_or_zero(_per_node(d1, l1))
    g2 = _inverse_or_zero(_per_node(d2, l2))
    i1 = l1.get_indexer(data[c1])
    i2 = l2.get_indexer(data[c2])
    n1, n2 = len(l1), len(l2)
    p12, x12, d12 = _build_csr(i1, i2, n1, n2, csr)
    p21, x21, d21 = _build_csr(i2, i1, n2, n1, csr)
    op12, op21 = HostOperator(n1, n2, p12, x12, g1), HostOperator(n2, n1, p21, x21, g2)
    op12.dev_csr, op21.dev_csr = d12, d21
    return set1, set2, list(l1), list(l2), op12, op21


def operator_from_edges(rows, cols, M, K, g=None) -> HostOperator:
    """Build an operator straight from index arrays (synthetic benchmarks, tests).
    ``g`` defaults to 1/row-degree (the unweighted reference normalisation)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    indptr, indices = _csr(rows, cols, M, K)
    deg = np.diff(indptr)
    if g is None:
        g = _inverse_or_zero(deg.astype(np.float64))
    return HostOperator(M, K, indptr, indices, np.asarray(g, dtype=np.float64), deg.astype(np.int64))
