"""Drop-in replacements for the classes of the reference's SimRank/SimRank.py, running the
iteration on a B200 through libsimrank_b200 (see simrank_b200/engine.py).

Same import surface (``from SimRank import SimRank`` then ``SimRank.SimRank()``), the same
no-argument constructors, ``fit`` signatures, return types, attributes and stdout text as the
reference (SimRank.py:8-493).  What differs is documented in DESIGN.md: graphs are kept as
CSR + row scale instead of dense DataFrames (``Graph``, ``Weight`` and ``Evidence`` are
materialised lazily on attribute access), bipartite results carry the labels that actually
belong to their rows by default, and every class offers ``top_k`` on the device-resident
result.  There is no CPU fallback: without a CUDA device ``fit`` raises.
"""
import sys
import time

import numpy as np
import pandas as pd

from SimRank.Helper import *  # noqa: F401,F403  (BAR_LENGTH, update_progress -- reference SimRank.py:6)
from SimRank.Helper import BAR_LENGTH, update_progress

from simrank_b200 import drivers as _drv


def _converged_message(iteration):
    return f'\rPercent: [{"#" * BAR_LENGTH}] 100% Complete! \n\rConverged at iteration {iteration}'


class _Base(object):
    """Engine options shared by every class.  All arguments are optional so that the
    reference's ``Cls()`` construction keeps working.

    mode   : None/'auto' | 'csr' (float64 gather path) | 'csr16' (uint16 fixed-point gather path) |
             'i8' (tcgen05 fixed-point dense path); 'auto' picks by size and density
    device : CUDA device (default: current)
    slices : uint8 planes per matrix in i8 mode: 2, 3, 4 or None/'auto' = the fewest planes whose
             guaranteed error bound stays below 5e-7 (simrank_b200.engine.choose_slices)
    sharded: None (default) | True | False.  With torch.distributed initialised and more than one
             rank, ``fit`` is a COLLECTIVE: every rank must call it with the same arguments, and S is
             row-sharded over the ranks' GPUs (one process per GPU).  ``sharded=False`` keeps the fit
             on the calling rank's own GPU; ``sharded=True`` insists on a process group.
    gather : row-sharded fits only: 'all' returns the whole matrix on every rank, 'local' returns
             each rank's own row block (all columns)
    result : 'frame' (the reference's DataFrames) or 'device' (fit returns the device-resident
             simrank_b200.drivers.Result; nothing is copied to the host until asked)
    label_order : bipartite only -- 'sorted' labels the result rows with the labels they belong
             to; 'reference' reproduces the set-ordered labels of SimRank.py:303
    """

    def _engine_options(self, mode=None, device=None, slices=None, label_order="sorted", gather="all",
                        result="frame", sharded=None):
        self._mode, self._device, self._slices, self._label_order = mode, device, slices, label_order
        self._sharded = sharded
        if result not in ("frame", "device"):
            raise ValueError("result must be 'frame' or 'device'")
        self._gather, self._result_kind = gather, result
        self.fit_info_ = None
        self._result = None

    def _tick(self, stage=None):
        """Wall-clock stages of the last fit (graph build, device setup, iterations, result
        transfer), kept in ``fit_timings_`` for the bench and for users who ask where time went."""
        now = time.perf_counter()
        if stage is None:
            self.fit_timings_ = {}
        else:
            self.fit_timings_[stage] = self.fit_timings_.get(stage, 0.0) + now - self._t_last
        self._t_last = now

    def _converged(self, s1, s2, eps):
        """True when no entry differs by more than ``eps`` (SimRank.py:54-77).  The engine fuses
        this reduction into the update epilogue; the method is kept for API compatibility."""
        return not bool((abs(np.asarray(s1) - np.asarray(s2)) > eps).sum())

    def _iterate(self, solver, iterations, eps, verbose, pair):
        if verbose:
            print('Start iterating...')
        on_it = (lambda it: update_progress(it / iterations)) if verbose else None
        halves = [h for h in (getattr(solver, "half", None), getattr(solver, "h1", None), getattr(solver, "h2", None))
                  if h is not None]
        initial = tuple(0.0 if h.n_out == 0 else 1.0 for h in halves)      # empty matrix: nothing differs
        applied, conv, last = _drv.run_loop(solver.step, iterations, eps, pair, on_it, initial)
        if conv and verbose:
            sys.stdout.write(_converged_message(applied))
            sys.stdout.flush()
        self.fit_info_ = _drv.FitInfo(applied, conv, last, solver.mode,
                                      tuple(tuple(getattr(h, "slices_used", ())) for h in halves),
                                      tuple(float(getattr(h, "err", 0.0)) for h in halves))
        return applied, conv

    def _finish(self, solver, labels):
        """Turn the finished solver into what ``fit`` returns: the reference's DataFrame(s)
        (SimRank.py:141, 303), or -- ``result="device"`` -- the device-resident Result itself
        (``.frame(i)``, ``.top_k(k, i)``, ``.mats[i]``), for matrices that must not be copied to the
        host as a whole (cfg5: S1 is 153 GB)."""
        self._result = _drv.collect(solver, labels, self._gather)
        if self._result_kind == "device":
            out = self._result
        elif len(labels) == 1:
            out = self._result.frame(0)
        else:
            out = tuple(self._result.frame(i) for i in range(len(labels)))
        self._tick("result")
        return out

    def top_k(self, k, which=0):
        """Row-wise top-k of the last fitted similarity matrix, computed on the device.
        Returns (labels DataFrame [n, k], values DataFrame [n, k]); ties resolve to the earlier
        column, exactly like ``np.argsort(-S, kind='stable')``."""
        if self._result is None:
            raise RuntimeError("call fit() first")
        return self._result.top_k(k, which)


# =========================================================================== directed
class SimRank(_Base):
    """SimRank on a weighted/unweighted directed graph (reference SimRank.py:8-141)."""

    def __init__(self, **engine_options):
        self.Nodes = set()
        self.Graph = pd.DataFrame()
        self._engine_options(**engine_options)

    # Graph is stored as an operator; the dense DataFrame of the reference (SimRank.py:43,52) is
    # only built when somebody reads the attribute.
    @property
    def Graph(self):
        if self._graph_frame is None and self._graph_op is not None:
            self._graph_frame = pd.DataFrame(self._graph_op.to_dense(), index=self._node_order,
                                             columns=self._node_order, copy=False)
        return self._graph_frame

    @Graph.setter
    def Graph(self, value):
        self._graph_frame, self._graph_op = value, None

    def _create_graph(self, data, weighted, from_node_column, to_node_column, weight_column):
        self.Nodes, self._node_order, self._graph_op = _drv.build_directed(
            data, weighted, from_node_column, to_node_column, weight_column)
        self._graph_frame = None

    def fit(self, data, C=0.8, weighted=False, from_node_column='from', to_node_column='to',
            weight_column='weight', iterations=100, eps=1e-4, verbose=True):
        self._tick()
        self._create_graph(data, weighted, from_node_column, to_node_column, weight_column)
        self._tick("graph")
        solver = _drv.directed_solver(self._graph_op, C, mode=self._mode, device=self._device, slices=self._slices,
                                      sharded=self._sharded)
        self._tick("setup")
        self._iterate(solver, iterations, eps, verbose, pair=False)
        self._tick("iterate")
        return self._finish(solver, [self._node_order])


class SimRankPP(SimRank):
    """SimRank++: evidence and spread weights (reference SimRank.py:305-363)."""

    def __init__(self, **engine_options):
        super(SimRankPP, self).__init__(**engine_options)
        self.Evidence = pd.DataFrame()
        self.Weight = pd.DataFrame()

    def _cal_Evidence(self, G, verbose):
        """Common-in-neighbour evidence ``1 - 0.5 ** (A A^T)`` (SimRank.py:311-320) as a lazy
        device object; ``np.asarray(obj)`` gives the reference's ndarray."""
        if verbose:
            print("Initializing Evidence matrix...")
        start = time.time()
        E = _drv.evidence(G, device=self._device)
        end = time.time()
        if verbose:
            print(f"Finished in {end - start}s!")
        return E

    def _cal_Weight(self, G, verbose):
        """``diag(exp(-var(row nonzeros))) G`` (SimRank.py:322-337) as an operator."""
        if verbose:
            print(f'Initializing Weight matrix...')
        start = time.time()
        W = _drv.weight(G, device=self._device)
        end = time.time()
        if verbose:
            print(f"Finished in {end - start}s!")
        return W

    def _prepare_pp(self, verbose):
        self.Weight = self._cal_Weight(self._graph_op, verbose)
        self.Evidence = self._cal_Evidence(self._graph_op, verbose)

    def fit(self, data, C=0.8, weighted=False, from_node_column='from', to_node_column='to',
            weight_column='weight', iterations=100, eps=1e-4, verbose=True):
        self._tick()
        self._create_graph(data, weighted, from_node_column, to_node_column, weight_column)
        self._prepare_pp(verbose)
        self._tick("graph")
        solver = _drv.directed_solver(self.Weight, C, evidence=self.Evidence, mode=self._mode,
                                      device=self._device, slices=self._slices, sharded=self._sharded)
        self._tick("setup")
        self._iterate(solver, iterations, eps, verbose, pair=False)
        self._tick("iterate")
        return self._finish(solver, [self._node_order])


class AprioriSimRank(SimRankPP):
    """SimRank++ blended with a prior similarity (reference SimRank.py:427-455)."""

    def __init__(self, **engine_options):
        super(AprioriSimRank, self).__init__(**engine_options)

    def fit(self, data, AprioriSim, C=0.8, lbd=0.5, weighted=False, from_node_column='from',
            to_node_column='to', weight_column='weight', iterations=100, eps=1e-4, verbose=True):
        self._tick()
        self._create_graph(data, weighted, from_node_column, to_node_column, weight_column)
        self._prepare_pp(verbose)
        self._tick("graph")
        solver = _drv.directed_solver(self.Weight, C, evidence=self.Evidence, prior=AprioriSim, lbd=lbd,
                                      mode=self._mode, device=self._device, slices=self._slices,
                                      sharded=self._sharded)
        self._tick("setup")
        self._iterate(solver, iterations, eps, verbose, pair=False)
        self._tick("iterate")
        return self._finish(solver, [self._node_order])


# =========================================================================== bipartite
class _BipartiteGraphs(object):
    """Lazy ``Graph_N1_N2`` / ``Graph_N2_N1`` frames (sorted-label order, SimRank.py:199-200)."""

    def _reset_graphs(self):
        self.NodesGroup1 = set()
        self.NodesGroup2 = set()
        self._op12 = self._op21 = None
        self._frames = {"12": pd.DataFrame(), "21": pd.DataFrame()}
        self._sorted1 = self._sorted2 = None

    def _frame(self, key):
        if self._frames[key] is None:
            op = self._op12 if key == "12" else self._op21
            idx, cols = (self._sorted1, self._sorted2) if key == "12" else (self._sorted2, self._sorted1)
            self._frames[key] = pd.DataFrame(op.to_dense(), index=idx, columns=cols, copy=False)
        return self._frames[key]

    Graph_N1_N2 = property(lambda self: self._frame("12"),
                           lambda self, v: self._frames.__setitem__("12", v))
    Graph_N2_N1 = property(lambda self: self._frame("21"),
                           lambda self, v: self._frames.__setitem__("21", v))

    def _create_graph(self, data, weighted, node_group1_column, node_group2_column, weight_column):
        (self.NodesGroup1, self.NodesGroup2, self._sorted1, self._sorted2,
         self._op12, self._op21) = _drv.build_bipartite(data, weighted, node_group1_column,
                                                        node_group2_column, weight_column)
        self._frames = {"12": None, "21": None}

    def _labels(self):
        if self._label_order == "reference":          # set order, as SimRank.py:303 labels them
            return list(self.NodesGroup1), list(self.NodesGroup2)
        return self._sorted1, self._sorted2


class BipartiteSimRank(_BipartiteGraphs, _Base):
    """SimRank on a bipartite graph (reference SimRank.py:143-303)."""

    def __init__(self, **engine_options):
        self._reset_graphs()
        self._engine_options(**engine_options)

    def fit(self, data, C1=0.8, C2=0.8, weighted=False, node_group1_column='user',
            node_group2_column='item', weight_column='weight', iterations=100, eps=1e-4, verbose=True):
        self._tick()
        self._create_graph(data, weighted, node_group1_column, node_group2_column, weight_column)
        self._tick("graph")
        solver = _drv.bipartite_solver(self._op12, self._op21, C1, C2, mode=self._mode, device=self._device,
                                       slices=self._slices, sharded=self._sharded)
        self._tick("setup")
        self._iterate(solver, iterations, eps, verbose, pair=True)
        self._tick("iterate")
        return self._finish(solver, list(self._labels()))


class BipartiteSimRankPP(_BipartiteGraphs, SimRankPP):
    """SimRank++ on a bipartite graph (reference SimRank.py:365-425).

    The reference multiplies BOTH updates by ``Evidence_N1`` (SimRank.py:423).  That is kept
    when the two groups have the same size (it is what the reference computes); when they
    differ the reference raises a broadcast error, and ``Evidence_N2`` is used instead."""

    def __init__(self, **engine_options):
        self._reset_graphs()
        self.Evidence_N1 = pd.DataFrame()
        self.Evidence_N2 = pd.DataFrame()
        self.Weight_N1 = pd.DataFrame()
        self.Weight_N2 = pd.DataFrame()
        self._engine_options(**engine_options)

    def _prepare_pp(self, verbose):
        self.Weight_N1 = self._cal_Weight(self._op12, verbose)
        self.Weight_N2 = self._cal_Weight(self._op21, verbose)
        self.Evidence_N1 = self._cal_Evidence(self._op12, verbose)
        self.Evidence_N2 = self._cal_Evidence(self._op21, verbose)

    def _group2_evidence(self):
        same = self._op12.M == self._op21.M
        return self.Evidence_N1 if same else self.Evidence_N2

    def fit(self, data, C1=0.8, C2=0.8, weighted=False, node_group1_column='user',
            node_group2_column='item', weight_column='weight', iterations=100, eps=1e-4, verbose=True):
        self._tick()
        self._create_graph(data, weighted, node_group1_column, node_group2_column, weight_column)
        self._prepare_pp(verbose)
        self._tick("graph")
        solver = _drv.bipartite_solver(self.Weight_N1, self.Weight_N2, C1, C2, evidence1=self.Evidence_N1,
                                       evidence2=self._group2_evidence(), mode=self._mode,
                                       device=self._device, slices=self._slices, sharded=self._sharded)
        self._tick("setup")
        self._iterate(solver, iterations, eps, verbose, pair=True)
        self._tick("iterate")
        return self._finish(solver, list(self._labels()))


class BipartitleAprioriSimRank(BipartiteSimRankPP):
    """Bipartite SimRank++ with priors (reference SimRank.py:457-493; the class name keeps the
    reference's spelling)."""

    def __init__(self, **engine_options):
        super(BipartitleAprioriSimRank, self).__init__(**engine_options)

    def fit(self, data, AprioriSim1, AprioriSim2, C1=0.8, C2=0.8, lbd1=0.5, lbd2=0.5, weighted=False,
            node_group1_column='user', node_group2_column='item', weight_column='weight',
            iterations=100, eps=1e-4, verbose=True):
        self._tick()
        self._create_graph(data, weighted, node_group1_column, node_group2_column, weight_column)
        self._prepare_pp(verbose)
        self._tick("graph")
        solver = _drv.bipartite_solver(self.Weight_N1, self.Weight_N2, C1, C2, evidence1=self.Evidence_N1,
                                       evidence2=self._group2_evidence(), prior1=AprioriSim1, prior2=AprioriSim2,
                                       lbd1=lbd1, lbd2=lbd2, mode=self._mode, device=self._device,
                                       slices=self._slices, sharded=self._sharded)
        self._tick("setup")
        self._iterate(solver, iterations, eps, verbose, pair=True)
        self._tick("iterate")
        return self._finish(solver, list(self._labels()))


# README.md:16 of the reference (and BASELINE.json) spell these "Bipartitle..."
BipartitleSimRank = BipartiteSimRank
BipartitleSimRankPP = BipartiteSimRankPP
