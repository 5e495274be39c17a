"""Row-sharded multi-GPU SimRank iteration (one process per GPU, torch.distributed / NCCL).

Partition.  Every similarity matrix is split in row blocks, rank r owning rows
``plan.start(r) .. plan.stop(r)``; the 0/1 adjacency pattern (dense uint8, 1 GB at n = 32768) is
replicated.  One update ``S_out <- epilogue(coef * G S_in G^T)`` is then

  1. MID   (local)   D[r, j] = sum_m S_in[r, m] A[j, m] for the LOCAL rows r of S_in and ALL j.
                     Because S_in is symmetric this is the column panel U[:, rows_r] of
                     U = A S_in; the kernel stores it transposed, one launch per destination
                     rank q, straight into the send block for q (rows_q of the panel).
  2. exchange        all-to-all of the uint8 planes: rank q receives U[rows_q, rows_r] from every
                     r, i.e. its ROW panel U[rows_q, :], K-blocked by source rank.
  3. FINAL (local)   S_out[rows_q, :] = epilogue(g g^T o (U[rows_q, :] A^T)) reading the receive
                     buffer in place through a 4-D tensor map (K-blocked operand).
  4. a 2-double MAX all-reduce gives every rank the same max|dS| (the reference's convergence
     test, SimRank.py:74) and the range of the new S.

The exchange moves NS bytes per element of U (3 with the default planes) instead of 8 for
float64.  The kernel launches are the same C-ABI calls as the single-GPU engine; the class is
written so that the two launch methods and the tensor device can be substituted, which is how
the world_size-2 gloo tests on CPU check every offset of the sharding (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import _ptr, _round_up, _stream
from .graph import HostOperator


class ShardPlan:
    """Row blocks of an n-row matrix over ``world`` ranks, padded to 128-row blocks for exchange."""

    def __init__(self, n: int, world: int):
        self.n, self.world = n, world
        self.per = -(-n // world) if n else 0
        self.blk = _round_up(max(self.per, 1), 128)

    def start(self, r):
        return min(self.n, r * self.per)

    def stop(self, r):
        return min(self.n, (r + 1) * self.per)

    def count(self, r):
        return self.stop(r) - self.start(r)

    @property
    def padded(self):
        return self.world * self.blk

    def pad_index(self, k: np.ndarray) -> np.ndarray:
        """Position of node k in the block-padded layout (block = owning rank)."""
        if self.per == 0:
            return k
        return (k // self.per) * self.blk + k % self.per


class ShardedHalf:
    """Rank-local state of one similarity matrix in the tensor-core (int8 planes) mode."""

    def __init__(self, op: HostOperator, coef, rank, world, device, ns=3, evidence=None, prior=None, lbd=0.0,
                 group=None):
        self.op, self.coef, self.rank, self.world, self.device, self.ns = op, float(coef), rank, world, device, ns
        self.group = group
        self.n_out, self.n_in = op.M, op.K
        self.out_plan, self.in_plan = ShardPlan(self.n_out, world), ShardPlan(self.n_in, world)
        self.row0, self.rows = self.out_plan.start(rank), self.out_plan.count(rank)
        self.ld = _round_up(max(self.n_out, 1), 16)
        self.ldp = _round_up(max(self.n_out, 1), 128)
        self.lda = max(_round_up(max(self.n_in, 1), 128), self.in_plan.padded)
        self.evidence, self.prior, self.lbd = evidence, prior, float(lbd)     # LOCAL rows
        self.events = None
        dev = device
        self.S = torch.zeros((max(self.rows, 1), self.ld), dtype=torch.float64, device=dev)
        self._init_identity()
        self.scal = torch.zeros(2, dtype=torch.float64, device=dev)
        self.maxoff = 0.0
        self.planes = torch.zeros((ns, max(self.rows, 1), self.ldp), dtype=torch.uint8, device=dev)
        bo, bi = self.out_plan.blk, self.in_plan.blk
        self.sendbuf = torch.zeros((world, ns, bo, bi), dtype=torch.uint8, device=dev)
        self.recvbuf = torch.zeros((world, ns, bo, bi), dtype=torch.uint8, device=dev)
        g = np.ascontiguousarray(op.g, dtype=np.float64)
        self.rho = g * op.deg
        self.rho_max = float(self.rho.max()) if self.rho.size else 0.0
        self.prior_max = float(prior.max()) if prior is not None else 0.0
        self.g = torch.from_numpy(g).to(dev)
        self.deg_dev = torch.from_numpy(op.deg.astype(np.float64)).to(dev)
        self.rho_dev = torch.from_numpy(np.ascontiguousarray(self.rho)).to(dev)
        self.bound_S, self.bound_S_max = (0.0, 1.0), 1.0
        # adjacency pattern, natural column layout (MID) and block-padded column layout (FINAL)
        self.a8_mid = self._dense_pattern(op.indices)
        padded_cols = self.in_plan.pad_index(op.indices.astype(np.int64)).astype(np.int32)
        same = bool(np.array_equal(padded_cols, op.indices))
        self.a8_fin = self.a8_mid if same else self._dense_pattern(padded_cols)

    # ---- device hooks (replaced by numpy stand-ins in the CPU tests) ------------------------
    def _init_identity(self):
        if self.rows:
            _lib.check(_lib.load().srk_set_identity_f64(_ptr(self.S), self.ld, self.rows, self.n_out, self.row0,
                                                        _stream()), "srk_set_identity_f64")

    def _dense_pattern(self, cols: np.ndarray) -> torch.Tensor:
        a8 = torch.empty((self.n_out, self.lda), dtype=torch.uint8, device=self.device)
        ptr = torch.from_numpy(self.op.indptr).to(self.device)
        idx = torch.from_numpy(np.ascontiguousarray(cols)).to(self.device)
        _lib.check(_lib.load().srk_csr_to_dense_u8(_ptr(ptr), _ptr(idx), 0, self.n_out, self.lda, _ptr(a8),
                                                   self.lda, _stream()), "srk_csr_to_dense_u8")
        return a8

    def _launch(self, args: _lib.I8Args, name: str):
        _lib.check(_lib.load().srk_i8_half(C.byref(args), _stream()), name)

    def _exchange(self):
        dist.all_to_all_single(self.recvbuf, self.sendbuf, group=self.group)

    def _reduce_scalars(self):
        dist.all_reduce(self.scal, op=dist.ReduceOp.MAX, group=self.group)

    # ---- one update ----------------------------------------------------------------------------
    def _timed(self, name, fn):
        if self.events is None:
            return fn()
        st = torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        self.events.append((name, a, b))

    def update(self, src: "ShardedHalf") -> None:
        ns = self.ns
        self.scal.zero_()
        guard = 1.0 + 2.0 ** -20
        s_off = min(src.bound_S_max, src.maxoff * (1.0 + 1e-6) + src.bound_S_max * 2.0 ** -23)
        u_mul, u_add = s_off * guard, guard
        blend = (1.0 - self.lbd) if self.prior is not None else 1.0
        mul = blend * self.coef * self.rho_max * max(1.0, s_off) * guard
        add = self.lbd * self.prior_max * guard if self.prior is not None else 0.0
        bo, bi = self.out_plan.blk, self.in_plan.blk

        def mids():
            for step in range(self.world):
                q = (self.rank + 1 + step) % self.world            # own block last
                nq = self.out_plan.count(q)
                if nq == 0 or src.rows == 0:
                    continue
                a = _lib.I8Args()
                a.mode, a.ns = _lib.SRK_I8_MID, ns
                a.R, a.N, a.K = src.rows, nq, self.n_in
                a.in_planes, a.ld_in, a.in_plane_stride = src.planes.data_ptr(), src.ldp, src.planes.stride(0)
                a.in_rowbound = _lib.RowBound.of(src.rho_dev.data_ptr() + 8 * src.row0, *src.bound_S)
                a.A8, a.lda = self.a8_mid.data_ptr() + self.out_plan.start(q) * self.lda, self.lda
                a.diag_offset, a.unit_diag = src.row0, 1
                a.out_planes = self.sendbuf.data_ptr() + q * ns * bo * bi
                a.ld_outp, a.out_plane_stride = bi, bo * bi
                a.out_rowbound = _lib.RowBound.of(self.deg_dev.data_ptr() + 8 * self.out_plan.start(q), u_mul, u_add)
                self._launch(a, "srk_i8_half(MID)")

        self._timed("i8_half_mid", mids)
        self._timed("exchange", self._exchange)

        def final():
            if self.rows == 0:
                return
            b = _lib.I8Args()
            b.mode, b.ns = _lib.SRK_I8_FINAL, ns
            b.R, b.N, b.K = self.rows, self.n_out, self.in_plan.padded
            b.in_planes, b.ld_in, b.in_plane_stride = self.recvbuf.data_ptr(), bi, bo * bi
            b.in_kblock, b.in_kblock_stride = bi, ns * bo * bi
            b.in_rowbound = _lib.RowBound.of(self.deg_dev.data_ptr() + 8 * self.row0, u_mul, u_add)
            b.A8, b.lda = self.a8_fin.data_ptr(), self.lda
            b.diag_offset, b.unit_diag = self.row0, 0
            b.g_row, b.g_col = self.g.data_ptr() + 8 * self.row0, self.g.data_ptr()
            b.out_f64, b.ld_out = self.S.data_ptr(), self.ld
            b.out_planes, b.ld_outp, b.out_plane_stride = self.planes.data_ptr(), self.ldp, self.planes.stride(0)
            b.out_rowbound = _lib.RowBound.of(self.rho_dev.data_ptr() + 8 * self.row0, mul, add)
            e = b.epi
            e.coef = self.coef
            if self.evidence is not None:
                e.evidence, e.ld_evidence = self.evidence.data_ptr(), self.evidence.stride(0)
            if self.prior is not None:
                e.prior, e.ld_prior, e.lambda_ = self.prior.data_ptr(), self.prior.stride(0), self.lbd
            e.s_old, e.ld_s_old = self.S.data_ptr(), self.ld
            e.maxdiff, e.maxoff = self.scal.data_ptr(), self.scal.data_ptr() + 8
            self._launch(b, "srk_i8_half(FINAL)")

        self._timed("i8_half_final", final)
        self._pending_bound = ((mul, add), mul * self.rho_max + add)

    def finish(self) -> float:
        self._reduce_scalars()
        maxdiff, maxoff = self.scal.tolist()
        self.maxoff = maxoff
        self.bound_S, self.bound_S_max = self._pending_bound
        return maxdiff

    def local_result(self) -> torch.Tensor:
        return self.S[: self.rows, : self.n_out]

    def gathered_result(self) -> torch.Tensor:
        """Full n_out x n_out matrix on every rank (all-gather of the row blocks)."""
        per = self.out_plan.per
        pad = torch.zeros((per, self.n_out), dtype=self.S.dtype, device=self.S.device)
        pad[: self.rows] = self.local_result()
        full = torch.empty((self.world * per, self.n_out), dtype=self.S.dtype, device=self.S.device)
        dist.all_gather_into_tensor(full, pad, group=self.group)
        return full[: self.n_out]


def _rank_world(group=None):
    return dist.get_rank(group), dist.get_world_size(group)


class ShardedDirectedSolver:
    """Row-sharded ``S <- [E o] C * W S W^T; diag <- 1`` (SimRank.py:139, :361)."""

    half_cls = ShardedHalf

    def __init__(self, op: HostOperator, C_, evidence=None, prior=None, lbd=0.0, mode="i8", ns=3, device=None,
                 group=None):
        if mode not in (None, "auto", "i8"):
            raise NotImplementedError("the row-sharded solver runs the tensor-core (i8) path")
        rank, world = _rank_world(group)
        self.mode = "i8"
        self.half = self.half_cls(op, C_, rank, world, device, ns, evidence, prior, lbd, group)
        self.halves = [self.half]

    def step(self) -> float:
        self.half.update(self.half)
        return self.half.finish()

    @property
    def S(self):
        return self.half.gathered_result()


class ShardedBipartiteSolver:
    """Row-sharded Gauss-Seidel alternation of SimRank.py:297-302."""

    half_cls = ShardedHalf

    def __init__(self, op12: HostOperator, op21: HostOperator, C1, C2, evidence1=None, evidence2=None, prior1=None,
                 prior2=None, lbd1=0.0, lbd2=0.0, mode="i8", ns=3, device=None, group=None):
        if mode not in (None, "auto", "i8"):
            raise NotImplementedError("the row-sharded solver runs the tensor-core (i8) path")
        rank, world = _rank_world(group)
        self.mode = "i8"
        self.h1 = self.half_cls(op12, C1, rank, world, device, ns, evidence1, prior1, lbd1, group)
        self.h2 = self.half_cls(op21, C2, rank, world, device, ns, evidence2, prior2, lbd2, group)
        self.halves = [self.h1, self.h2]

    def step(self):
        self.h1.update(self.h2)
        d1 = self.h1.finish()
        self.h2.update(self.h1)
        d2 = self.h2.finish()
        return d1, d2

    @property
    def S1(self):
        return self.h1.gathered_result()

    @property
    def S2(self):
        return self.h2.gathered_result()
