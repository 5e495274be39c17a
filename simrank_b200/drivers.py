"""Glue between the drop-in classes (SimRank/SimRank.py) and the device engine."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import engine as _eng
from .engine import FitInfo, run_loop  # noqa: F401  (re-exported for SimRank/SimRank.py)
from . import graph as _graph
from .graph import HostOperator


def build_directed(data, weighted, from_node_column, to_node_column, weight_column):
    """graph.build_directed with the CSR arrays built on the GPU (engine.device_csr)."""
    return _graph.build_directed(data, weighted, from_node_column, to_node_column, weight_column, csr=_eng.device_csr)


def build_bipartite(data, weighted, node_group1_column, node_group2_column, weight_column):
    """graph.build_bipartite with the CSR arrays built on the GPU (engine.device_csr)."""
    return _graph.build_bipartite(data, weighted, node_group1_column, node_group2_column, weight_column,
                                  csr=_eng.device_csr)


def _device_op(op: HostOperator, device=None) -> _eng.DeviceOperator:
    dev = _eng.require_cuda(device)
    cached = getattr(op, "_dev", None)
    if cached is None or cached.device != dev:
        cached = _eng.DeviceOperator(op, dev)
        op._dev = cached
    return cached


class WeightOperator(HostOperator):
    """SimRank++ weight matrix ``W = diag(spread) * G`` kept as an operator.  ``np.asarray(W)``
    materialises the dense ndarray the reference stores in ``self.Weight`` (SimRank.py:333)."""

    def __array__(self, dtype=None, copy=None):
        out = self.to_dense()
        return out if dtype is None else out.astype(dtype)

    @property
    def shape(self):
        return (self.M, self.K)


def weight(G: HostOperator, device=None) -> WeightOperator:
    """``_cal_Weight`` (SimRank.py:322-337): spread from the row-variance kernel, then an O(n)
    rescale of the row factors instead of the reference's n^3 ``np.dot(spread, G)``."""
    dop = _device_op(G, device)
    spread = dop.row_spread().cpu().numpy()
    W = WeightOperator(G.M, G.K, G.indptr, G.indices, G.g * spread, G.deg)
    W._dev = dop.with_scale(spread)
    W.spread = spread
    return W


class EvidenceMatrix:
    """Evidence ``1 - 0.5 ** (A A^T)`` (SimRank.py:315-316) held on the device as uint8
    common-neighbour counts (a count >= 54 already gives exactly 1.0 in float64), computed on
    first use: the tensor-core path reads the evidence of an operator's own pattern from the
    uint16 ``A A^T`` counts it needs anyway.  ``np.asarray(E)`` materialises the float64 ndarray
    of the reference."""

    def __init__(self, G: HostOperator, device=None, mode: str = "auto"):
        self.op, self.n, self._device, self._mode, self._counts = G, G.M, device, mode, None

    @property
    def counts(self) -> torch.Tensor:
        if self._counts is None:
            self._counts = _device_op(self.op, self._device).evidence_counts(self._mode)
        return self._counts

    def is_pattern_of(self, op: HostOperator) -> bool:
        """True when this evidence was computed from the 0/1 pattern of ``op`` (W = diag(spread) G
        shares the pattern of G) and no row is dead (g <= 0 rows count as empty, SimRank.py:315)."""
        same = self.op.indices is op.indices and self.op.indptr is op.indptr and self.op.M == op.M
        return bool(same and not (self.op.dead.astype(bool) & (self.op.deg > 0)).any())

    @property
    def shape(self):
        return (self.n, self.n)

    def __array__(self, dtype=None, copy=None):
        cnt = self.counts[:, : self.n].cpu().numpy().astype(np.int64)
        out = 1 - 0.5 ** cnt
        return out if dtype is None else out.astype(dtype)


def evidence(G: HostOperator, device=None, mode: str = "auto") -> EvidenceMatrix:
    _eng.require_cuda(device)
    return EvidenceMatrix(G, device, mode)


def _prior_tensor(prior, n, device):
    if prior is None:
        return None
    arr = np.ascontiguousarray(np.asarray(prior, dtype=np.float64))
    if arr.shape != (n, n):
        raise ValueError(f"prior must have shape {(n, n)}, got {arr.shape}")
    return torch.from_numpy(arr).to(device)


def _mode_with_prior(mode, *priors):
    # the fixed-point path needs S >= 0; a negative prior entry can break that
    if any(p is not None and np.asarray(p).min() < 0 for p in priors):
        return "csr"
    return mode


def _require_symmetric_prior(*priors):
    """The row-sharded solvers read column blocks of S as transposed row blocks, which is only the
    same thing while S stays symmetric: it does for every class of the reference unless an Apriori
    prior is not symmetric itself (SimRank.py:453)."""
    for p in priors:
        if p is None:
            continue
        a = np.asarray(p, dtype=np.float64)
        if a.ndim != 2 or a.shape[0] != a.shape[1] or float(np.abs(a - a.T).max(initial=0.0)) > 1e-12 * max(
                1.0, float(np.abs(a).max(initial=0.0))):
            raise NotImplementedError("a row-sharded fit needs a symmetric prior matrix; run a non-symmetric "
                                      "AprioriSim on one GPU")


def _world(sharded=None):
    """(rank, world) a fit runs on.  ``sharded=None`` (default): the default process group when
    torch.distributed is initialised with more than one rank -- every rank must then call ``fit``
    with the same arguments (it is a collective: S is row-sharded over the ranks) -- else (0, 1).
    ``sharded=False`` keeps a fit on the calling rank's own GPU whatever the process group;
    ``sharded=True`` insists on a process group; a ``dist.LocalRank`` runs the fit as one of P logical
    ranks that share one GPU (dist.LocalCluster)."""
    import torch.distributed as dist
    from .dist import LocalRank
    if isinstance(sharded, LocalRank):                # a logical rank of a dist.LocalCluster (one GPU)
        return sharded.rank, sharded.world
    active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if sharded is False or (sharded is None and not active):
        return 0, 1
    if not active:
        raise RuntimeError("sharded=True needs an initialised torch.distributed process group with world size > 1")
    return dist.get_rank(), dist.get_world_size()


def _group_of(sharded):
    from .dist import LocalRank
    return sharded if isinstance(sharded, LocalRank) else None


def _local_rows(t, n, rank, world):
    """Rows of an n-row device matrix owned by ``rank`` under the row sharding of dist.ShardPlan."""
    if t is None:
        return None
    from .dist import ShardPlan
    plan = ShardPlan(n, world)
    return t[plan.start(rank):plan.stop(rank)]


def _evidence_args(evidence, op: HostOperator, mode: str):
    """(uint8 counts tensor or None, take-it-from-the-pattern flag) for one update."""
    if evidence is None:
        return None, False
    if mode in ("i8", "csr16") and evidence.is_pattern_of(op):
        return None, True
    return evidence.counts, False


def directed_solver(op: HostOperator, C, evidence=None, prior=None, lbd=0.0, mode=None, device=None, slices=None,
                    sharded=None):
    dop = _device_op(op, device)
    pr = _prior_tensor(prior, op.M, dop.device)
    rank, world = _world(sharded)
    if world > 1:                      # one process per GPU: S row-sharded
        from . import dist as _sd
        _require_symmetric_prior(prior)
        smode = _sd._sharded_mode(_mode_with_prior(mode, prior), op, coefs=(C,), lbds=(lbd,),
                                  has_prior=prior is not None)
        ev, from_pattern = _evidence_args(evidence, op, smode)
        return _sd.ShardedDirectedSolver(op, C, _local_rows(ev, op.M, rank, world),
                                         _local_rows(pr, op.M, rank, world), lbd, smode,
                                         slices, dop.device, group=_group_of(sharded),
                                         evidence_from_pattern=from_pattern)
    mode = _eng.choose_mode(op, _mode_with_prior(mode, prior), C, lbd, prior is not None)
    ev, from_pattern = _evidence_args(evidence, op, mode)
    return _eng.DirectedSolver(dop, C, ev, pr, lbd, mode, slices,
                               evidence_from_pattern=from_pattern)


def bipartite_solver(op12: HostOperator, op21: HostOperator, C1, C2, evidence1=None, evidence2=None, prior1=None,
                     prior2=None, lbd1=0.0, lbd2=0.0, mode=None, device=None, slices=None, sharded=None):
    d12, d21 = _device_op(op12, device), _device_op(op21, device)
    p1 = _prior_tensor(prior1, op12.M, d12.device)
    p2 = _prior_tensor(prior2, op21.M, d21.device)
    rank, world = _world(sharded)
    if world > 1:
        from . import dist as _sd
        _require_symmetric_prior(prior1, prior2)
        smode = _sd._sharded_mode(_mode_with_prior(mode, prior1, prior2), op12, op21, coefs=(C1, C2),
                                  lbds=(lbd1, lbd2), has_prior=prior1 is not None or prior2 is not None)
        e1, pat1 = _evidence_args(evidence1, op12, smode)
        e2, pat2 = _evidence_args(evidence2, op21, smode)
        return _sd.ShardedBipartiteSolver(op12, op21, C1, C2, _local_rows(e1, op12.M, rank, world),
                                          _local_rows(e2, op21.M, rank, world), _local_rows(p1, op12.M, rank, world),
                                          _local_rows(p2, op21.M, rank, world), lbd1, lbd2,
                                          smode, slices, d12.device, group=_group_of(sharded),
                                          evidence1_from_pattern=pat1, evidence2_from_pattern=pat2)
    mode = _mode_with_prior(mode, prior1, prior2)
    m1 = _eng.choose_mode(op12, mode, C1, lbd1, prior1 is not None)
    m2 = _eng.choose_mode(op21, mode, C2, lbd2, prior2 is not None)
    mode = m1 if m1 == m2 else "csr"
    e1, pat1 = _evidence_args(evidence1, op12, mode)
    e2, pat2 = _evidence_args(evidence2, op21, mode)
    return _eng.BipartiteSolver(d12, d21, C1, C2, e1, e2, p1, p2, lbd1, lbd2, mode, slices,
                                evidence1_from_pattern=pat1, evidence2_from_pattern=pat2)


class _PinnedPool:
    """Page-locked host buffers for result transfers, reused across fits.  Page-locking a fresh
    gigabyte costs as much as copying it, so a buffer whose last user is gone (the DataFrame built on
    it has been dropped: a weak reference to the ndarray the frame was built on tells) is
    handed out again; a buffer that is still referenced is never reused.  Bounded by
    SIMRANK_B200_PINNED_POOL_BYTES (default 16 GiB, so that the 8.6 GB result of BASELINE cfg4 on one GPU
    is kept: page-locking it afresh cost 0.03 - 0.13 s of a 0.9 s fit); larger results get a fresh
    allocation that is not kept."""

    def __init__(self):
        self.entries = []                                      # [base tensor, weakref to the ndarray handed out]

    def take(self, shape, dtype):
        import os
        import weakref
        n = int(np.prod(shape))
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if nbytes < (1 << 20):
            t = torch.empty(shape, dtype=dtype)                # small: pageable is fine
            return t, t.numpy()
        for e in self.entries:
            if e[0].numel() == n and e[0].dtype == dtype and (e[1] is None or e[1]() is None):
                view = e[0].view(shape)
                arr = view.numpy()
                e[1] = weakref.ref(arr)
                return view, arr
        base = torch.empty(n, dtype=dtype, pin_memory=True)
        view = base.view(shape)
        arr = view.numpy()
        if nbytes <= int(os.environ.get("SIMRANK_B200_PINNED_POOL_BYTES", 16 << 30)):
            self.entries = self.entries[-3:] + [[base, weakref.ref(arr)]]
        return view, arr


_PINNED = _PinnedPool()


class Result:
    """Device-resident result of a fit: similarity matrices + their labels.  ``rows[w]`` is the
    (first, last+1) range of rows of matrix ``w`` held here: everything on a single GPU or with
    ``gather="all"``, the local row block of a row-sharded fit with ``gather="local"``."""

    def __init__(self, mats, labels, rows=None):
        self.mats, self.labels = mats, labels
        self.rows = rows if rows is not None else [(0, len(lab)) for lab in labels]

    def _row_labels(self, which):
        lo, hi = self.rows[which]
        lab = self.labels[which]
        return lab if (lo, hi) == (0, len(lab)) else lab[lo:hi]

    def frame(self, which: int = 0) -> pd.DataFrame:
        """The labelled DataFrame the reference returns (SimRank.py:141, 303)."""
        S = self.mats[which]
        host, arr = _PINNED.take(tuple(S.shape), S.dtype)
        host.copy_(S)
        return pd.DataFrame(arr, index=self._row_labels(which), columns=self.labels[which], copy=False)

    def top_k(self, k: int, which: int = 0):
        S = self.mats[which]
        idx, vals = _eng.topk_rows(S, k)
        lab = np.asarray(self.labels[which], dtype=object)
        idx_h = idx.cpu().numpy()
        rows = self._row_labels(which)
        return (pd.DataFrame(lab[idx_h], index=rows), pd.DataFrame(vals.cpu().numpy(), index=rows))


def collect(solver, labels, gather: str = "all") -> Result:
    """Result of a finished solver.  A row-sharded solver either all-gathers every matrix onto
    every rank (``gather="all"``: each rank returns what the reference returns) or keeps the local
    row blocks (``gather="local"``: rank r returns rows ``plan.start(r):plan.stop(r)`` of each
    matrix, nothing crosses NVLink or PCIe twice)."""
    halves = getattr(solver, "halves", None)
    pair = hasattr(type(solver), "S1")          # on the class: the property getter all-gathers the matrix
    if halves is None or gather == "all":
        return Result([solver.S1, solver.S2] if pair else [solver.S], labels)
    if gather != "local":
        raise ValueError("gather must be 'all' or 'local'")
    return Result([h.local_result() for h in halves], labels, [(h.row0, h.row0 + h.rows) for h in halves])
