"""Row-sharded multi-GPU SimRank iteration (one process per GPU, torch.distributed / NCCL).

Partition.  Every similarity matrix is split in row blocks, rank q owning rows
``plan.start(q) .. plan.stop(q)`` (blocks are multiples of 16 rows); the 0/1 adjacency pattern
(dense uint8, 1 GB at n = 32768) is replicated.  One update ``S_out <- epilogue(coef * G S_in G^T)``
of the tensor-core path (engine.py, ``S = I + S_off``) is, on rank q:

  1. slice   planes of the LOCAL rows of S_in with their exact row maxima (srk_slice_rows_max_f64).
  2. MID     for every destination rank p:  U[j, r] = sum_k A[j, k] S_off[r, k] for j in rows_p and
             the local rows r.  Because S_in is symmetric this is the column block (rows_p x rows_q)
             of U = A S_off, and the kernel stores it straight into rank p's ROW panel U[rows_p, :]
             at column offset start(q) -- through a peer-mapped pointer over NVLink (symmetric
             memory): the exchange is fused into the epilogue of the GEMM, there is no all-to-all.
  3. barrier every rank's row panel of U is complete.
  4. FINAL   S_out[rows_q, :] from the own row panel of U.  S_out is symmetric, so each unordered
             pair of row blocks {q, p} is computed ONCE: rank q computes its diagonal block with the
             symmetric layout and the blocks (q, p) for p up to half-way round the ring with the
             transposed layout, whose epilogue also writes the mirror image into rank p's S_out
             (peer pointer again).  With an even world the blocks half-way round are split in two.
             Every rank so executes 1/P of the n^3 flop of the single-GPU symmetric FINAL.
  5. a 2-double MAX all-reduce gives every rank the same max|dS| (the reference's convergence
     test, SimRank.py:74) and the range of the new S; it is also the barrier that orders the
     mirror stores before the next update.

``StagedExchange`` is the same algorithm without peer pointers: the kernels store into local
staging blocks that torch.distributed moves (all-to-all) -- the fallback when symmetric memory is
not available and the way the host logic runs under gloo on CPU (tests/test_dist_gloo.py, with the
numpy emulator of the C ABI).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import (ListSplit, _ptr, _round_up, _stream, attach_sync_ws, choose_slices, count_bits, gather_qmax,
                     slice_delta)
from .graph import HostOperator

_NO_DIAGONAL = -(1 << 40)


class ShardPlan:
    """Row blocks of an n-row matrix over ``world`` ranks; every block starts at a multiple of 16
    (the kernels store 16-element chunks at the block's column offset)."""

    def __init__(self, n: int, world: int):
        self.n, self.world = n, world
        self.per = _round_up(-(-n // world), 16) if n else 0

    def start(self, r):
        return min(self.n, r * self.per)

    def stop(self, r):
        return min(self.n, (r + 1) * self.per)

    def count(self, r):
        return self.stop(r) - self.start(r)


def pair_tasks(plan: ShardPlan, q: int, symmetric: bool):
    """Blocks of the symmetric result that rank ``q`` computes in FINAL, as tuples
    (p, j_lo, j_hi, r_lo, r_hi, mirror): A8 rows j_lo..j_hi of rank p's block against the local
    rows r_lo..r_hi; ``mirror`` says whether the transposed block also goes to rank p.  The own
    block (p == q) is computed with the symmetric layout.  Without symmetry (a prior) every rank
    computes all its rows."""
    P = plan.world
    rows_q = plan.count(q)
    tasks = []
    if rows_q == 0:
        return tasks
    if not symmetric:
        return [(p, 0, plan.count(p), 0, rows_q, False) for p in range(P) if plan.count(p)]
    tasks.append((q, 0, rows_q, 0, rows_q, False))
    for d in range(1, P):
        p = (q + d) % P
        rows_p = plan.count(p)
        if rows_p == 0:
            continue
        if 2 * d < P:
            tasks.append((p, 0, rows_p, 0, rows_q, True))
        elif 2 * d == P:
            # the pair meets half-way round the ring in both directions: split the block by the
            # rows of the HIGHER rank
            hi, lo = max(q, p), min(q, p)
            h = min(_round_up(plan.count(hi) // 2, 16), plan.count(hi))
            if q == lo:
                if h > 0:
                    tasks.append((p, 0, h, 0, rows_q, True))
            elif h < rows_q:
                tasks.append((p, 0, rows_p, h, rows_q, True))
    return tasks


# --------------------------------------------------------------------------- communicators
# Every collective of the sharded solvers goes through these four helpers, so that ``group`` can be a
# torch.distributed process group (None = the default one) or a LocalRank: one of P logical ranks that
# share ONE GPU inside one process (LocalCluster below).
def _rank_world(group=None):
    if isinstance(group, LocalRank):
        return group.rank, group.world
    return dist.get_rank(group), dist.get_world_size(group)


def _all_reduce_max(t, group):
    if isinstance(group, LocalRank):
        return group.all_reduce_max(t)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)


def _all_to_all(recv, send, group):
    if isinstance(group, LocalRank):
        return group.all_to_all(recv, send)
    if recv.dtype == torch.int16:                      # NCCL has no 16-bit integer type: move the bytes
        recv, send = recv.view(torch.uint8), send.view(torch.uint8)
    dist.all_to_all_single(recv, send, group=group)


def _all_gather_into(full, part, group):
    if isinstance(group, LocalRank):
        return group.all_gather_into(full, part)
    dist.all_gather_into_tensor(full, part, group=group)


class LocalCluster:
    """P logical ranks on ONE GPU in one process, one Python thread per rank (SURVEY.md section 4 iv:
    "same kernels, P logical shards on 1 GPU").  The ranks run the unmodified sharded solvers; their
    collectives are thread barriers plus device copies, and "peer" pointers are simply the other
    ranks' buffers on the same device, so the kernels' peer-store paths (U blocks, mirrored S blocks,
    row-maximum keys) execute exactly as they do over NVLink.  All ranks launch on the same CUDA
    stream, and a launch is enqueued before its thread reaches the next barrier, so stream order
    respects every barrier.  Used by the single-GPU tests of the multi-GPU algorithms and for
    debugging a sharded fit without a multi-GPU box."""

    def __init__(self, world: int):
        import threading
        self.world = world
        self.barrier = threading.Barrier(world)
        self.lock = threading.Lock()
        self.slots = {}
        self.arenas = {}

    def run(self, fn):
        """Call ``fn(group)`` on every logical rank (group = its LocalRank); -> list of results."""
        import threading
        out, err = [None] * self.world, [None] * self.world
        dev = torch.cuda.current_device() if torch.cuda.is_available() else None

        def body(r):
            try:
                if dev is not None:
                    torch.cuda.set_device(dev)
                out[r] = fn(LocalRank(self, r))
            except BaseException as exc:                       # noqa: BLE001 -- re-raised below
                err[r] = exc
                self.barrier.abort()

        threads = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        real = [e for e in err if e is not None and not isinstance(e, __import__("threading").BrokenBarrierError)]
        if real or any(err):
            raise (real or [e for e in err if e is not None])[0]
        return out


class LocalRank:
    """Communicator handle of one logical rank of a LocalCluster (passed as ``group``)."""

    def __init__(self, cluster: LocalCluster, rank: int):
        self.cluster, self.rank, self.world = cluster, rank, cluster.world
        self._allocs = 0

    def _publish(self, key, value):
        c = self.cluster
        with c.lock:
            c.slots.setdefault(key, [None] * self.world)[self.rank] = value
        c.barrier.wait()
        return c.slots[key]

    def _done(self):
        self.cluster.barrier.wait()                            # everybody has read the published values

    def all_reduce_max(self, t):
        parts = self._publish("max", t)
        m = torch.stack([x.to(t.device) for x in parts]).amax(dim=0)
        self._done()
        t.copy_(m)

    def all_to_all(self, recv, send):
        parts = self._publish("a2a", send)
        for s in range(self.world):
            recv[s].copy_(parts[s][self.rank])
        self._done()

    def all_gather_into(self, full, part):
        parts = self._publish("gather", part)
        n = part.shape[0]
        for s in range(self.world):
            full[s * n:(s + 1) * n].copy_(parts[s])
        self._done()

    def barrier(self):
        self.cluster.barrier.wait()

    def alloc_shared(self, shape, dtype, device):
        """The rank's buffer of a symmetric allocation and the device pointers of every rank's."""
        c, key = self.cluster, self._allocs
        self._allocs += 1
        with c.lock:
            if key not in c.arenas:
                c.arenas[key] = [torch.zeros(shape, dtype=dtype, device=device) for _ in range(self.world)]
        bufs = c.arenas[key]
        return bufs[self.rank], [int(b.data_ptr()) for b in bufs]


# --------------------------------------------------------------------------- exchange strategies
class StagedExchange:
    """Kernels store into local staging blocks; collectives move them."""

    peer = False

    def __init__(self, group=None):
        self.group = group

    def alloc(self, shape, dtype, device):
        return torch.zeros(shape, dtype=dtype, device=device), None

    def barrier(self):
        pass


class PeerExchange:
    """Symmetric memory: every rank maps the others' buffers; kernels store into them directly."""

    peer = True

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group if group is not None else dist.group.WORLD
        self._handles = []

    def alloc(self, shape, dtype, device):
        t = self.symm.empty(*shape, dtype=dtype, device=device)
        t.zero_()
        h = self.symm.rendezvous(t, self.group)
        self._handles.append(h)
        return t, [int(x) for x in h.buffer_ptrs]

    def barrier(self):
        self._handles[0].barrier(channel=0)


class LocalPeerExchange:
    """PeerExchange of a LocalCluster: the peers' buffers live on the same device."""

    peer = True

    def __init__(self, group: LocalRank):
        self.group = group

    def alloc(self, shape, dtype, device):
        return self.group.alloc_shared(tuple(shape), dtype, device)

    def barrier(self):
        self.group.barrier()


_PEER_PROBE = {}                 # (device, group) -> symmetric memory works on every rank (probed once per process)


def make_exchange(device, group=None):
    """Peer-memory exchange when every rank can map symmetric memory, else the staged fallback.  The
    decision is agreed across the ranks (MIN all-reduce of a probe allocation + rendezvous), so that a
    box without P2P / NVLink takes the staged path on all ranks instead of failing on some; the probe
    (a rendezvous: ~0.5 s) runs once per process and group."""
    want = os.environ.get("SIMRANK_B200_EXCHANGE", "auto").lower()
    if isinstance(group, LocalRank):
        return StagedExchange(group) if want == "staged" else LocalPeerExchange(group)
    if want == "staged" or torch.device(device).type != "cuda":
        return StagedExchange(group)
    key = (str(device), id(group) if group is not None else None)
    if key in _PEER_PROBE:
        if _PEER_PROBE[key]:
            return PeerExchange(group)
        if want == "peer":
            raise RuntimeError("SIMRANK_B200_EXCHANGE=peer, but symmetric memory is not available on every rank")
        return StagedExchange(group)
    ex, ok = None, 1
    try:
        ex = PeerExchange(group)
        ex.symm.empty(256, dtype=torch.uint8, device=device)          # local part of the probe
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()):
        try:
            ex.alloc((256,), torch.uint8, device)                      # collective part: rendezvous
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    _PEER_PROBE[key] = bool(int(flag.item()))
    if _PEER_PROBE[key]:
        ex._handles.clear()                                            # barrier() uses the first REAL allocation
        return ex
    if want == "peer":
        raise RuntimeError("SIMRANK_B200_EXCHANGE=peer, but symmetric memory is not available on every rank")
    return StagedExchange(group)


class ShardedHalf:
    """Rank-local state of one similarity matrix in the tensor-core (paired-SM) mode."""

    def __init__(self, op: HostOperator, coef, rank, world, device, ns=None, evidence=None, prior=None, lbd=0.0,
                 group=None, evidence_from_pattern=False, exchange=None):
        self.op, self.coef, self.rank, self.world, self.device, self.ns = op, float(coef), rank, world, device, ns
        self.group = group
        self.n_out, self.n_in = op.M, op.K
        self.plan = ShardPlan(self.n_out, world)
        self.row0, self.rows, self.per = self.plan.start(rank), self.plan.count(rank), max(self.plan.per, 16)
        self.ld = _round_up(max(self.n_out, 1), 16)                 # S, counts: columns of the output
        self.ldp = _round_up(max(self.n_out, 1), 128)               # planes of S_out (used as a source)
        self.ldu = _round_up(max(self.n_in, 1), 128)                # row panel of U
        self.lda = _round_up(max(self.n_in, 1), 128)
        self.evidence, self.prior, self.lbd = evidence, prior, float(lbd)     # LOCAL rows
        self.evidence_from_pattern = bool(evidence_from_pattern)
        self.symmetric = prior is None
        self.events = None
        self.slices_used = []
        self.err = self._err_next = 0.0                             # see engine._Half.err
        self.ns_alloc = 3 if ns in (None, "auto") else int(ns)
        self.ex = exchange if exchange is not None else make_exchange(device, group)
        dev = device
        # S, the row panel of U and the row-maximum keys are written by other ranks in peer mode: one
        # peer-mapped arena per matrix (a symmetric-memory rendezvous costs ~0.1 s whatever its size)
        s_bytes = self.per * self.ld * 8
        u_bytes = self.ns_alloc * self.per * self.ldu
        k_bytes = _round_up(self.per * 4, 256)
        arena, ptrs = self.ex.alloc((_round_up(s_bytes, 256) + _round_up(u_bytes, 256) + k_bytes,), torch.uint8, dev)
        o_u, o_k = _round_up(s_bytes, 256), _round_up(s_bytes, 256) + _round_up(u_bytes, 256)
        self.S = arena[:s_bytes].view(torch.float64).view(self.per, self.ld)
        self.U = arena[o_u:o_u + u_bytes].view(self.ns_alloc, self.per, self.ldu)
        self.S_ptrs = ptrs and [q for q in ptrs]
        self.U_ptrs = ptrs and [q + o_u for q in ptrs]
        self._init_identity()
        self.planes = None
        # keys of the row maxima of the local rows of S, collected by the FINAL epilogues of every rank
        # that writes into them (peer mode; the staged fallback takes the maxima in the slicer)
        self.rowmax = arena[o_k:o_k + self.per * 4].view(torch.int32) if self.ex.peer else None
        self.rowmax_ptrs = ptrs and [q + o_k for q in ptrs]
        self.bound_vec = torch.zeros(self.per, dtype=torch.float64, device=dev)
        self.scal = torch.zeros(2, dtype=torch.float64, device=dev)
        self.maxoff = 0.0
        g = np.ascontiguousarray(op.g, dtype=np.float64)
        self.rho_max = float((g * op.deg).max()) if g.size else 0.0
        self.g = torch.from_numpy(g).to(dev)
        self.deg_dev = torch.from_numpy(op.deg.astype(np.float64)).to(dev)
        self.a8 = self._dense_pattern()
        self.counts = self._pattern_counts()
        self.tasks = pair_tasks(self.plan, rank, self.symmetric)
        if not self.ex.peer:
            # staging: U blocks per destination rank, mirror blocks per partner rank
            self.send_U = torch.zeros((world, self.ns_alloc, self.per, 16), dtype=torch.uint8, device=dev)
            self.mirror_send = torch.zeros((world, self.per, self.per), dtype=torch.float64, device=dev) \
                if self.symmetric and world > 1 else None
            self.mirror_recv = torch.zeros_like(self.mirror_send) if self.mirror_send is not None else None
        self.version, self._sliced = 0, (-1, 0)

    # ---- device hooks (replaced by numpy stand-ins in the CPU tests) ------------------------
    def _init_identity(self):
        if self.rows:
            _lib.check(_lib.load().srk_set_identity_f64(_ptr(self.S), self.ld, self.rows, self.n_out, self.row0,
                                                        _stream()), "srk_set_identity_f64")

    def _dense_pattern(self) -> torch.Tensor:
        a8 = torch.empty((self.n_out, self.lda), dtype=torch.uint8, device=self.device)
        ptr = torch.from_numpy(self.op.indptr).to(self.device)
        idx = torch.from_numpy(np.ascontiguousarray(self.op.indices)).to(self.device)
        _lib.check(_lib.load().srk_csr_to_dense_u8(_ptr(ptr), _ptr(idx), 0, self.n_out, self.lda, _ptr(a8),
                                                   self.lda, _stream()), "srk_csr_to_dense_u8")
        return a8

    def _launch(self, args: _lib.X2Args, name: str):
        attach_sync_ws(args, self.device)
        _lib.check(_lib.load().srk_x2_half(C.byref(args), _stream()), name)

    def _launch_slice(self, ns: int):
        if self.rowmax is not None:                               # one pass: the maxima are known
            _lib.check(_lib.load().srk_slice_rows_key_f64(
                _ptr(self.S), self.ld, self.rows, self.n_out, self.row0, ns, _ptr(self.rowmax), _ptr(self.planes),
                self.ldp, self.planes.stride(0), _ptr(self.bound_vec), _stream()), "srk_slice_rows_key_f64")
            return
        _lib.check(_lib.load().srk_slice_rows_max_f64(
            _ptr(self.S), self.ld, self.rows, self.n_out, self.row0, ns, _ptr(self.planes), self.ldp,
            self.planes.stride(0), _ptr(self.bound_vec), _stream()), "srk_slice_rows_max_f64")

    def _reduce_scalars(self):
        _all_reduce_max(self.scal, self.group)

    # ---- pieces of one update ------------------------------------------------------------------
    def _pattern_counts(self) -> torch.Tensor:
        """uint16 ``A A^T`` for the LOCAL rows (rows_q x n_out), see DeviceOperator.pattern_counts."""
        ldc = _round_up(max(self.n_out, 1), 16)
        bits = count_bits(self.op.deg)
        cnt = torch.zeros((self.per, ldc), dtype=torch.int16 if bits == 16 else torch.int32, device=self.device)
        if self.rows:
            a = _lib.X2Args()
            a.mode, a.ns = _lib.SRK_X2_COUNTS, 1
            a.M, a.R, a.K = self.rows, self.n_out, self.n_in
            a.A8, a.lda = self.a8.data_ptr() + self.row0 * self.lda, self.lda
            a.in_planes, a.ld_in, a.in_plane_stride = self.a8.data_ptr(), self.lda, self.a8.numel()
            a.out_counts, a.ld_out_counts, a.counts_bits = cnt.data_ptr(), ldc, bits
            self._launch(a, "srk_x2_half(COUNTS)")
        return cnt

    def _timed(self, name, fn):
        if self.events is None:
            return fn()
        st = torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        self.events.append((name, a, b))

    def _planes_for(self, ns: int) -> torch.Tensor:
        if self.planes is None or self.planes.shape[0] < ns:
            self.planes = torch.zeros((max(ns, self.ns_alloc), self.per, self.ldp), dtype=torch.uint8,
                                      device=self.device)
            self._sliced = (-1, 0)
        if self._sliced != (self.version, ns):
            if self.rows:
                self._timed("slice_rows_key" if self.rowmax is not None else "slice_rows_max", lambda: self._launch_slice(ns))
            self._sliced = (self.version, ns)
        return self.planes

    def _mid(self, src: "ShardedHalf", ns: int, bound_mul: float):
        """Column block (rows_p x src rows) of U for every destination rank p."""
        planes_in = src._planes_for(ns)
        if not self.ex.peer and self.send_U.shape[-1] != src.per:
            self.send_U = torch.zeros((self.world, self.ns_alloc, self.per, src.per), dtype=torch.uint8,
                                      device=self.device)
        if src.rows == 0:
            return
        for step in range(self.world):
            p = (self.rank + 1 + step) % self.world               # own block last
            rows_p = self.plan.count(p)
            if rows_p == 0:
                continue
            a = _lib.X2Args()
            a.mode, a.ns = _lib.SRK_X2_MID, ns
            a.M, a.R, a.K = rows_p, src.rows, self.n_in
            a.A8, a.lda = self.a8.data_ptr() + self.plan.start(p) * self.lda, self.lda
            a.in_planes, a.ld_in, a.in_plane_stride = planes_in.data_ptr(), src.ldp, planes_in.stride(0)
            a.in_rowbound = _lib.RowBound.of(src.bound_vec.data_ptr(), 1.0, 0.0)
            if self.ex.peer:                                      # rank p's row panel, my columns
                a.out_planes = self.U_ptrs[p] + src.row0
                a.ld_outp, a.out_plane_stride = self.ldu, self.per * self.ldu
            else:
                a.out_planes = self.send_U[p].data_ptr()
                a.ld_outp, a.out_plane_stride = src.per, self.per * src.per
            a.out_rowbound = _lib.RowBound.of(self.deg_dev.data_ptr() + 8 * self.plan.start(p), bound_mul, 0.0)
            self._launch(a, "srk_x2_half(MID)")

    def _exchange_U(self, src: "ShardedHalf", ns: int):
        if self.ex.peer:
            self.ex.barrier()
            return
        recv = torch.empty_like(self.send_U)
        _all_to_all(recv, self.send_U, self.group)
        for s in range(self.world):                               # block from rank s = its rows as my columns
            cnt = src.plan.count(s)
            if cnt:
                self.U[:ns, :, src.plan.start(s):src.plan.start(s) + cnt] = recv[s, :ns, :, :cnt]

    def _final(self, ns: int, bound_mul: float):
        for (p, j_lo, j_hi, r_lo, r_hi, mirror) in self.tasks:
            b = _lib.X2Args()
            b.mode, b.ns = _lib.SRK_X2_FINAL, ns
            own = p == self.rank and self.symmetric
            col0 = self.plan.start(p) + j_lo                       # first output column of this block
            b.M, b.R, b.K = j_hi - j_lo, r_hi - r_lo, self.n_in
            b.A8, b.lda = self.a8.data_ptr() + col0 * self.lda, self.lda
            b.in_planes = self.U.data_ptr() + r_lo * self.ldu
            b.ld_in, b.in_plane_stride = self.ldu, self.per * self.ldu
            b.in_rowbound = _lib.RowBound.of(self.deg_dev.data_ptr() + 8 * (self.row0 + r_lo), bound_mul, 0.0)
            b.g_a, b.g_v = self.g.data_ptr() + 8 * col0, self.g.data_ptr() + 8 * (self.row0 + r_lo)
            off = r_lo * self.ld + col0                            # element offset of the block inside S / counts
            ldc = self.counts.stride(0)
            esz = self.counts.element_size()
            b.counts, b.ld_counts, b.add_counts = self.counts.data_ptr() + esz * (r_lo * ldc + col0), ldc, 1
            b.counts_bits = 8 * esz
            b.use_evidence = 1 if self.evidence_from_pattern else 0
            b.out_f64, b.ld_out = self.S.data_ptr() + 8 * off, self.ld
            e = b.epi
            e.coef = self.coef
            e.s_old, e.ld_s_old = self.S.data_ptr() + 8 * off, self.ld
            e.maxdiff, e.maxoff = self.scal.data_ptr(), self.scal.data_ptr() + 8
            if self.evidence is not None and not self.evidence_from_pattern:
                e.evidence = self.evidence.data_ptr() + r_lo * self.evidence.stride(0) + col0
                e.ld_evidence = self.evidence.stride(0)
            if self.prior is not None:
                e.prior = self.prior.data_ptr() + 8 * (r_lo * self.prior.stride(0) + col0)
                e.ld_prior, e.lambda_ = self.prior.stride(0), self.lbd
            if self.rowmax is not None:
                b.rowmax_hi = self.rowmax.data_ptr() + 4 * r_lo    # own block: r_lo == j_lo == 0
            if own:
                b.layout, b.diag_offset = _lib.SRK_X2_SYMMETRIC, 0
            else:
                b.layout = _lib.SRK_X2_TRANSPOSED
                # the diagonal runs through the own block only: j + col0 == row0 + r_lo + r
                b.diag_offset = (self.row0 + r_lo - col0) if p == self.rank else _NO_DIAGONAL
                if mirror:
                    if self.ex.peer:                               # rows j of rank p's S, my columns
                        b.mirror_out = self.S_ptrs[p] + 8 * (j_lo * self.ld)
                        b.ld_mirror, b.mirror_col0 = self.ld, self.row0 + r_lo
                        b.mirror_rowmax_hi = self.rowmax_ptrs[p] + 4 * j_lo
                    else:
                        b.mirror_out = self.mirror_send[p].data_ptr() + 8 * (j_lo * self.per)
                        b.ld_mirror, b.mirror_col0 = self.per, r_lo
            self._launch(b, "srk_x2_half(FINAL)")

    def _exchange_mirrors(self):
        """Staged mode: deliver the mirrored blocks and place them (peer mode stored them already)."""
        if self.ex.peer or self.mirror_send is None:
            return
        _all_to_all(self.mirror_recv, self.mirror_send, self.group)
        for s in range(self.world):
            if s == self.rank:
                continue
            for (p, j_lo, j_hi, r_lo, r_hi, mirror) in pair_tasks(self.plan, s, True):
                if p == self.rank and mirror:                      # rank s computed rows j_lo..j_hi of mine
                    c0 = self.plan.start(s) + r_lo
                    self.S[j_lo:j_hi, c0:c0 + (r_hi - r_lo)] = self.mirror_recv[s, j_lo:j_hi, r_lo:r_hi]

    # ---- one update ----------------------------------------------------------------------------
    def update(self, src: "ShardedHalf") -> None:
        self.scal.zero_()
        blend = (1.0 - self.lbd) if self.prior is not None else 1.0
        ns = choose_slices(self.ns, self.coef, blend, self.rho_max, src.maxoff)
        if ns > self.U.shape[0]:
            # every rank takes this branch together: ns derives from the all-reduced range of S_in
            self.ns_alloc = ns
            self.U, self.U_ptrs = self.ex.alloc((ns, self.per, self.ldu), torch.uint8, self.device)
            if not self.ex.peer:
                self.send_U = torch.zeros((self.world, ns, self.per, 16), dtype=torch.uint8, device=self.device)
        self.slices_used.append(ns)
        self._err_next = blend * self.coef * self.rho_max ** 2 * src.err + \
            slice_delta(ns, self.coef, blend, self.rho_max, src.maxoff)
        guard = 1.0 + 2.0 ** -14
        bound_mul = src.maxoff * guard                             # U[j, :] <= deg_j * max(S_off)
        self._timed("x2_half_mid", lambda: self._mid(src, ns, bound_mul))
        if self.rowmax is not None:
            self.rowmax.zero_()          # after the slice inside _mid read the old keys, before the barrier
        self._timed("exchange", lambda: self._exchange_U(src, ns))
        self._timed("x2_half_final", lambda: self._final(ns, bound_mul))
        self._timed("mirror_exchange", self._exchange_mirrors)
        self.version += 1

    def finish(self) -> float:
        self._reduce_scalars()
        maxdiff, maxoff = self.scal.tolist()
        self.maxoff = maxoff
        self.err = self._err_next
        return maxdiff

    def local_result(self) -> torch.Tensor:
        return self.S[: self.rows, : self.n_out]

    def gathered_result(self) -> torch.Tensor:
        """Full n_out x n_out matrix on every rank (all-gather of the row blocks)."""
        per = self.per
        pad = torch.zeros((per, self.n_out), dtype=self.S.dtype, device=self.S.device)
        pad[: self.rows] = self.local_result()
        full = torch.empty((self.world * per, self.n_out), dtype=self.S.dtype, device=self.S.device)
        _all_gather_into(full, pad, self.group)
        return full[: self.n_out]


class ShardedCsrHalf:
    """Rank-local state of one similarity matrix in the float64 CSR mode (exact arithmetic; graphs the
    fixed-point path cannot hold: negative weight sums, or when float64 results are asked for).

    S_out is row-sharded like in ShardedHalf.  One update ``S_out <- epilogue(coef * G S_in G^T)`` on
    rank q, with T = (G S_in)^T (n_in x n_out):

      1. first half   T[rows_in_q, :]: ``srk_csr_half_f64`` on X = (local rows of S_in)^T -- S_in is
                      symmetric, so the local ROW block transposed is the COLUMN block the gather needs.
                      One launch per destination rank p over the graph rows of p's block, stored
                      straight into the send block for p.
      2. exchange     block transpose: rank p receives T[rows_in_q, rows_out_p] from every q
                      (``all_to_all_single``) and so holds the column panel
                      T[:, rows_out_p] as an [n_in x rows_out_p] matrix.
      3. second half  ``srk_csr_half_f64`` with the fused epilogue on that panel: S_out[rows_out_p, :],
                      ``diag_offset`` = first global row of the block.
      4. the 2-double MAX all-reduce of ShardedHalf.finish.
    """

    def __init__(self, op: HostOperator, coef, rank, world, device, evidence=None, prior=None, lbd=0.0, group=None):
        self.op, self.coef, self.rank, self.world, self.device, self.group = op, float(coef), rank, world, device, group
        self.n_out, self.n_in = op.M, op.K
        self.plan = ShardPlan(self.n_out, world)
        self.row0, self.rows, self.per = self.plan.start(rank), self.plan.count(rank), max(self.plan.per, 16)
        self.ld = _round_up(max(self.n_out, 1), 16)
        self.evidence, self.prior, self.lbd = evidence, prior, float(lbd)          # LOCAL rows
        self.events = None
        self.slices_used = []
        self.S = torch.zeros((self.per, self.ld), dtype=torch.float64, device=device)
        self._init_identity()
        self.scal = torch.zeros(2, dtype=torch.float64, device=device)
        self.maxoff = 0.0
        self.indptr = torch.from_numpy(op.indptr).to(device)
        self.indices = torch.from_numpy(np.ascontiguousarray(op.indices)).to(device)
        self.g = torch.from_numpy(np.ascontiguousarray(op.g, dtype=np.float64)).to(device)
        self._send = self._recv = None

    # ---- device hooks (replaced by numpy stand-ins in the CPU tests) ------------------------
    def _init_identity(self):
        if self.rows:
            _lib.check(_lib.load().srk_set_identity_f64(_ptr(self.S), self.ld, self.rows, self.n_out, self.row0,
                                                        _stream()), "srk_set_identity_f64")

    def _launch_csr(self, row_begin, row_end, x_ptr, ldx, L, out_ptr, ldo, epi):
        import ctypes as C_
        _lib.check(_lib.load().srk_csr_half_f64(_ptr(self.indptr), _ptr(self.indices), _ptr(self.g), self.n_out,
                                                row_begin, row_end, C_.c_void_p(x_ptr), ldx, L, C_.c_void_p(out_ptr),
                                                ldo, C_.byref(epi) if epi is not None else None, _stream()),
                   "srk_csr_half_f64")

    def _reduce_scalars(self):
        _all_reduce_max(self.scal, self.group)

    def _library(self):
        """The C ABI (replaced by the numpy emulator in the CPU tests)."""
        return _lib.load()

    _timed = ShardedHalf._timed

    # ---- one update ----------------------------------------------------------------------------
    def update(self, src: "ShardedCsrHalf") -> None:
        self.scal.zero_()
        P, per_in, per_out = self.world, src.per, self.per
        if self._send is None or self._send.shape != (P, per_in, per_out):
            self._send = torch.zeros((P, per_in, per_out), dtype=torch.float64, device=self.device)
            self._recv = torch.zeros((P, per_in, per_out), dtype=torch.float64, device=self.device)

        def first():
            if src.rows == 0:
                return
            # column block of the symmetric S_in = its local row block, transposed (a copy, no arithmetic);
            # rows padded to the block height so that they stay 16-byte aligned for the TMA gather
            xt = torch.empty((self.n_in, per_in), dtype=torch.float64, device=self.device)
            xt[:, : src.rows] = src.S[: src.rows, : self.n_in].t()
            for p in range(P):
                lo, hi = self.plan.start(p), self.plan.stop(p)
                if hi > lo:        # OUT[c, i] lands at send[p][c, i - lo]: shift the base by -lo columns
                    self._launch_csr(lo, hi, xt.data_ptr(), per_in, src.rows,
                                     self._send[p].data_ptr() - 8 * lo, per_out, None)
        self._timed("csr_half_first", first)
        self._timed("exchange", lambda: _all_to_all(self._recv, self._send, self.group))

        def second():
            if self.rows == 0:
                return
            e = _lib.Epilogue()
            e.coef = self.coef
            if self.evidence is not None:
                e.evidence, e.ld_evidence = self.evidence.data_ptr(), self.evidence.stride(0)
            if self.prior is not None:
                e.prior, e.ld_prior, e.lambda_ = self.prior.data_ptr(), self.prior.stride(0), self.lbd
            e.s_old, e.ld_s_old = self.S.data_ptr(), self.ld
            e.maxdiff, e.maxoff = self.scal.data_ptr(), self.scal.data_ptr() + 8
            e.diag_offset = self.row0
            # block q of the receive buffer holds rows start_in(q).. of the panel: [P * per_in, per_out]
            if getattr(self, "evidence_from_pattern", False):       # a float64 update of the csr16 mode
                b = _lib.CsrArgs()
                b.elem, b.mode = _lib.SRK_ELEM_F64, _lib.SRK_CSR_FINAL
                b.indptr, b.indices, b.g = self.indptr.data_ptr(), self.indices.data_ptr(), self.g.data_ptr()
                b.M, b.row_begin, b.row_end = self.n_out, 0, self.n_out
                b.X, b.ldx, b.L, b.K = self._recv.data_ptr(), per_out, self.rows, P * per_in
                b.OUT, b.ldo = self.S.data_ptr(), self.ld
                b.counts, b.ld_counts = self.counts.data_ptr(), self.counts.stride(0)
                b.counts_bits, b.use_evidence = 8 * self.counts.element_size(), 1
                b.epi = e
                _lib.check(self._library().srk_csr_half(C.byref(b), _stream()), "srk_csr_half(f64, second)")
                return
            self._launch_csr(0, self.n_out, self._recv.data_ptr(), per_out, self.rows, self.S.data_ptr(), self.ld, e)
        self._timed("csr_half_final", second)

    def finish(self) -> float:
        self._reduce_scalars()
        maxdiff, maxoff = self.scal.tolist()
        self.maxoff = maxoff
        self.err = getattr(self, "_err_next", 0.0)
        return maxdiff

    local_result = ShardedHalf.local_result
    gathered_result = ShardedHalf.gathered_result


class ShardedCsr16Half(ShardedCsrHalf):
    """ShardedCsrHalf with the gathers in uint16 fixed point (engine._Half csr16 mode, row-sharded):

      1. quantise  Xq = srk_quantize_rows_u16(local rows of S_in): the transposed local row block with one
                   unit per local row -- the column block of S_off the first half gathers from.
      2. first     Tq[rows_in_q, rows_out_p] per destination rank p (exact integer sums, re-quantised
                   with the bound deg(i) * max(S_off) of its column), stored into the send block for p.
      3. exchange  the same block transpose, 2 bytes per element instead of 8.
      4. second    srk_csr_half FINAL on the received panel with counts = A A^T of the local rows (the
                   unit-diagonal term, also the SimRank++ evidence) and the fused epilogue.

    An update whose 16-bit error bound would exceed engine.ERR_BUDGET runs in float64 (the parent's
    update): both kinds read and write the same float64 S."""

    def __init__(self, op: HostOperator, coef, rank, world, device, evidence=None, prior=None, lbd=0.0, group=None,
                 evidence_from_pattern=False):
        super().__init__(op, coef, rank, world, device, evidence, prior, lbd, group)
        if prior is not None:
            raise ValueError("mode='csr16' does not take a prior; use mode='csr' (float64)")
        self.evidence_from_pattern = bool(evidence_from_pattern)
        g = np.ascontiguousarray(op.g, dtype=np.float64)
        self.rho_max = float((g * op.deg).max()) if g.size else 0.0
        self.deg_dev = torch.from_numpy(op.deg.astype(np.float64)).to(device)
        self.qmax = gather_qmax(op.deg)                            # of the gathers THIS half runs
        self.lda = _round_up(max(self.n_in, 1), 128)
        self.a8 = self._dense_pattern()
        self.counts = self._pattern_counts()
        del self.a8                                                # only needed for the counts
        self.err = self._err_next = 0.0
        self.ldxt = _round_up(self.per, 64)
        self.Xq, self.unit = None, torch.zeros(self.per, dtype=torch.float64, device=device)
        self.version, self._quantized_version = 0, (-1, 0.0)
        self._send16 = self._recv16 = None
        # hub rows (the popular items of a ratings graph) are pre-summed in pieces, see engine.ListSplit
        self.split = ListSplit.plan(self.indptr, self.indices, self.n_in, int(op.deg.max()) if op.deg.size else 0)
        # The second half runs as SRK_CSR_ACCUM over EVERY list + SRK_CSR_FINISH (include/simrank_b200.h): the
        # gather launch has no epilogue and no transposition tile (15.9 TB/s on BASELINE cfg5's S1 shape, the
        # L2 roof, against 9.0 for the fused launch whose 16-row CTAs idle through their epilogues), the
        # epilogue streams afterwards: 76.8 -> 60.0 ms there, 2.95 -> 2.14 ms on one rank of cfg4 over 8 GPUs
        # (profiles/r2_csr_shapes_finish.jsonl).  SRK_FINAL_VIA_ACCUM=0 keeps the fused launch.
        self.split_all = None
        if os.environ.get("SRK_FINAL_VIA_ACCUM", "1") == "1" and op.nnz:
            self.split_all = ListSplit.plan(self.indptr, self.indices, self.n_in, int(op.deg.max()), all_rows=True)
        # The first half goes the same way (ACCUM over every list, then SRK_CSR_FINISH_FIRST per destination
        # block) when a 1 KB-wide panel of its operand stays in L2; when it does not (S2 of cfg5: 142 MB) most
        # lists span several ranges and every piece would be an atomic add: there the hub split + the fused
        # launch stay.  SRK_FIRST_VIA_ACCUM=0 keeps the fused launch everywhere.
        self.first_via_accum = (self.split_all is not None and self.n_in * 1024 <= 64 * 2 ** 20 and
                                os.environ.get("SRK_FIRST_VIA_ACCUM", "1") == "1")

    _dense_pattern = ShardedHalf._dense_pattern
    _pattern_counts = ShardedHalf._pattern_counts
    _launch = ShardedHalf._launch

    def _quantized(self, qmax: float = 65535.0):
        """uint16 transposed local row block of the CURRENT S (cached per version and range; ``qmax`` is
        the consumer's, engine.gather_qmax)."""
        if self.Xq is None:
            self.Xq = torch.zeros((self.n_out, self.ldxt), dtype=torch.int16, device=self.device)
        if self._quantized_version != (self.version, qmax) and self.rows:
            _lib.check(self._library().srk_quantize_rows_u16(_ptr(self.S), self.ld, self.rows, self.n_out, self.row0,
                                                             _ptr(self.Xq), self.ldxt, _ptr(self.unit), qmax, 0, _stream()),
                       "srk_quantize_rows_u16")
        self._quantized_version = (self.version, qmax)
        return self.Xq, self.unit

    def _args(self, elem, mode):
        a = _lib.CsrArgs()
        a.elem, a.mode = elem, mode
        a.indptr, a.indices, a.g = self.indptr.data_ptr(), self.indices.data_ptr(), self.g.data_ptr()
        a.M = self.n_out
        return a

    def update(self, src: "ShardedCsr16Half") -> None:
        kappa = self.coef * self.rho_max ** 2
        qmax = self.qmax
        if getattr(self, "force_f64", False) or choose_slices(None, self.coef, 1.0, self.rho_max,
                                                              src.maxoff * 65536.0 / (qmax + 1.0)) > 2:
            self.slices_used.append(0)                              # 0 = float64 update
            self._err_next = kappa * src.err
            super().update(src)
            self.version += 1
            return
        self.slices_used.append(2)
        self._err_next = kappa * src.err + slice_delta(2, self.coef, 1.0, self.rho_max, src.maxoff) * 65536.0 / (qmax + 1.0)
        self.scal.zero_()
        lib = self._library()
        P, per_in, per_out = self.world, src.per, self.per
        if self._send16 is None or self._send16.shape != (P, per_in, per_out):
            self._send16 = torch.zeros((P, per_in, per_out), dtype=torch.int16, device=self.device)
            self._recv16 = torch.zeros((P, per_in, per_out), dtype=torch.int16, device=self.device)
        guard = 1.0 + 2.0 ** -14
        bound_mul = src.maxoff * guard                              # U[i, :] <= deg_i * max(S_off)

        def first():
            if src.rows == 0:
                return
            xq, unit = src._quantized(qmax)
            via = self.split_all if self.first_via_accum else None
            if via is not None:                                     # one gather launch for every graph row
                via.accumulate(lib, self.indices.data_ptr(), xq.data_ptr(), src.ldxt, src.rows, self.n_in, qmax)
            elif self.split is not None:                            # one launch for the hub rows of every block
                self.split.accumulate(lib, self.indices.data_ptr(), xq.data_ptr(), src.ldxt, src.rows, self.n_in, qmax)
            for p in range(P):
                lo, hi = self.plan.start(p), self.plan.stop(p)
                if hi <= lo:
                    continue
                a = self._args(_lib.SRK_ELEM_U16, _lib.SRK_CSR_FIRST)
                a.row_begin, a.row_end = lo, hi
                a.X, a.ldx, a.L, a.K = xq.data_ptr(), src.ldxt, src.rows, self.n_in
                a.OUT, a.ldo = self._send16[p].data_ptr() - 2 * lo, per_out      # column i lands at i - lo
                a.in_unit = _lib.RowBound.of(unit.data_ptr(), 1.0, 0.0)
                a.out_bound = _lib.RowBound.of(self.deg_dev.data_ptr(), bound_mul, 0.0)
                a.qmax = qmax
                if via is not None:
                    a.mode, a.accum, a.ld_accum = _lib.SRK_CSR_FINISH_FIRST, via._accum.data_ptr(), via._accum.shape[1]
                elif self.split is not None:
                    self.split.attach(a)
                _lib.check(lib.srk_csr_half(C.byref(a), _stream()), "srk_csr_half(u16, first)")
        self._timed("csr16_half_first", first)
        self._timed("exchange", lambda: _all_to_all(self._recv16, self._send16, self.group))

        def second():
            if self.rows == 0:
                return
            b = self._args(_lib.SRK_ELEM_U16, _lib.SRK_CSR_FINAL)
            b.row_begin, b.row_end = 0, self.n_out
            b.X, b.ldx, b.L, b.K = self._recv16.data_ptr(), per_out, self.rows, P * per_in
            b.OUT, b.ldo = self.S.data_ptr(), self.ld
            b.in_unit = _lib.RowBound.of(self.deg_dev.data_ptr() + 8 * self.row0, bound_mul / qmax, 0.0)
            b.qmax = qmax
            b.g_col = self.g.data_ptr() + 8 * self.row0
            esz = self.counts.element_size()
            b.counts, b.ld_counts, b.counts_bits, b.add_counts = self.counts.data_ptr(), self.counts.stride(0), 8 * esz, 1
            b.use_evidence = 1 if self.evidence_from_pattern else 0
            e = b.epi
            e.coef = self.coef
            if self.evidence is not None and not self.evidence_from_pattern:
                e.evidence, e.ld_evidence = self.evidence.data_ptr(), self.evidence.stride(0)
            e.s_old, e.ld_s_old = self.S.data_ptr(), self.ld
            e.maxdiff, e.maxoff = self.scal.data_ptr(), self.scal.data_ptr() + 8
            e.diag_offset = self.row0
            if self.split_all is not None:
                self.split_all.accumulate(lib, self.indices.data_ptr(), b.X, b.ldx, b.L, b.K, qmax)
                b.mode = _lib.SRK_CSR_FINISH
                b.accum, b.ld_accum = self.split_all._accum.data_ptr(), self.split_all._accum.shape[1]
            elif self.split is not None:
                self.split.accumulate(lib, self.indices.data_ptr(), b.X, b.ldx, b.L, b.K, qmax)
                self.split.attach(b)
            _lib.check(lib.srk_csr_half(C.byref(b), _stream()), "srk_csr_half(u16, second)")
        self._timed("csr16_half_final", second)
        self.version += 1


def _sharded_mode(mode, *ops, coefs=None, lbds=None, has_prior=False) -> str:
    """'i8' (tensor-core planes, peer-memory exchange) or 'csr' (float64, all-to-all exchange), by the
    rules of engine.choose_mode so that 'auto' picks the same arithmetic on 1 and on N GPUs: the
    fixed-point planes need finite, non-negative row scales, C > 0 and 0 <= lbd <= 1 (asked for
    explicitly on anything else it fails loudly), and 'auto' keeps small, very sparse or
    non-contracting problems on the float64 path."""
    from .engine import choose_mode
    mode = (mode or "auto").lower()
    if mode not in ("auto", "i8", "csr", "csr16"):
        raise ValueError(f"unknown mode {mode!r} for the row-sharded solver")
    coefs = coefs or (0.8,) * len(ops)
    lbds = lbds or (0.0,) * len(ops)
    if torch.cuda.is_available():
        picks = {choose_mode(op, mode, c, l, has_prior) for op, c, l in zip(ops, coefs, lbds)}
    else:            # CPU emulation of the host logic (tests/test_dist_gloo.py): no device to ask
        from .engine import fixed_point_obstacle
        bad = [fixed_point_obstacle(op, c, l, has_prior) for op, c, l in zip(ops, coefs, lbds)]
        if mode in ("i8", "csr16") and any(bad):
            raise ValueError(f"mode={mode!r} {next(b for b in bad if b)}; use mode='csr' (float64)")
        picks = {mode if mode in ("csr", "csr16") else ("csr" if b else "i8") for b in bad}
    return picks.pop() if len(picks) == 1 else "csr"


class ShardedDirectedSolver:
    """Row-sharded ``S <- [E o] C * W S W^T; diag <- 1`` (SimRank.py:139, :361)."""

    half_cls = ShardedHalf
    csr_half_cls = ShardedCsrHalf
    csr16_half_cls = ShardedCsr16Half

    def __init__(self, op: HostOperator, C_, evidence=None, prior=None, lbd=0.0, mode="i8", ns=None, device=None,
                 group=None, evidence_from_pattern=False):
        rank, world = _rank_world(group)
        self.mode = _sharded_mode(mode, op, coefs=(C_,), lbds=(lbd,), has_prior=prior is not None)
        if self.mode == "csr16":
            self.half = self.csr16_half_cls(op, C_, rank, world, device, evidence, prior, lbd, group,
                                            evidence_from_pattern)
        elif self.mode == "csr":
            if evidence_from_pattern:
                raise ValueError("the CSR path takes evidence counts, not the pattern flag")
            self.half = self.csr_half_cls(op, C_, rank, world, device, evidence, prior, lbd, group)
        else:
            self.half = self.half_cls(op, C_, rank, world, device, ns, evidence, prior, lbd, group,
                                      evidence_from_pattern)
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
    csr_half_cls = ShardedCsrHalf
    csr16_half_cls = ShardedCsr16Half

    def __init__(self, op12: HostOperator, op21: HostOperator, C1, C2, evidence1=None, evidence2=None, prior1=None,
                 prior2=None, lbd1=0.0, lbd2=0.0, mode="i8", ns=None, device=None, group=None,
                 evidence1_from_pattern=False, evidence2_from_pattern=False):
        rank, world = _rank_world(group)
        self.mode = _sharded_mode(mode, op12, op21, coefs=(C1, C2), lbds=(lbd1, lbd2),
                                  has_prior=prior1 is not None or prior2 is not None)
        if self.mode == "csr16":
            self.h1 = self.csr16_half_cls(op12, C1, rank, world, device, evidence1, prior1, lbd1, group,
                                          evidence1_from_pattern)
            self.h2 = self.csr16_half_cls(op21, C2, rank, world, device, evidence2, prior2, lbd2, group,
                                          evidence2_from_pattern)
            self.halves = [self.h1, self.h2]
            return
        if self.mode == "csr":
            if evidence1_from_pattern or evidence2_from_pattern:
                raise ValueError("the CSR path takes evidence counts, not the pattern flag")
            self.h1 = self.csr_half_cls(op12, C1, rank, world, device, evidence1, prior1, lbd1, group)
            self.h2 = self.csr_half_cls(op21, C2, rank, world, device, evidence2, prior2, lbd2, group)
            self.halves = [self.h1, self.h2]
            return
        self.h1 = self.half_cls(op12, C1, rank, world, device, ns, evidence1, prior1, lbd1, group,
                                evidence1_from_pattern)
        self.h2 = self.half_cls(op21, C2, rank, world, device, ns, evidence2, prior2, lbd2, group,
                                evidence2_from_pattern, exchange=self.h1.ex if self.h1.ex.peer else None)
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
