"""Device-side SimRank iteration engine (single GPU; the row-sharded variant is in dist.py).

Host orchestration of the kernels behind include/simrank_b200.h.  PyTorch owns device
memory and streams; every numeric step is a hand-written sm_100a kernel reached through
the C ABI -- there is no torch/numpy arithmetic on the hot path and no CPU fallback.

One update of the reference loop (SimRank.py:138-140)

    old_S = deepcopy(new_S); new_S = C * G.dot(new_S).dot(G.T); fill_diagonal(new_S, 1)

plus the reduction behind ``_converged`` (SimRank.py:74) is two kernel launches here:

  csr mode   T = (G S)^T                     srk_csr_half_f64   (gather, transposed store)
             S = epilogue((G T)^T)           srk_csr_half_f64   (fused epilogue, in place)
  csr16 mode Xq = uint16(S_off), column units srk_quantize_rows_u16  (S = I + S_off)
             Tq = uint16((A S_off)^T)        srk_csr_half FIRST  (integer gather, re-quantised)
             S = epilogue(g g^T o (A Tq + A A^T))
                                             srk_csr_half FINAL  (integer gather, fused epilogue, in
                                                                 place; pairs r >= i only, mirrored)
  i8 mode    planes(S_off), exact row bounds srk_slice_rows_max_f64   (S = I + S_off)
             U = A S_off (re-quantised)      srk_x2_half MID    (tcgen05 cta_group::2 kind::i8)
             S = epilogue(g g^T o (A U^T + A A^T))
                                             srk_x2_half FINAL  (fused epilogue, in place; only
                                                                 the upper triangle is computed,
                                                                 the lower one is mirrored)

S is updated in place: each epilogue thread reads S_old[r, c] for max|dS| and then writes
S_new[r, c]; nothing else reads S during the second half.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .graph import HostOperator

_NS_DEFAULT = None          # None/'auto': fewest planes whose a-priori error bound meets ERR_BUDGET
# Guaranteed max-abs deviation from the float64 iteration caused by the fixed-point planes
# (the north-star tolerance is 1e-6; the float64 epilogue itself differs by ~1e-16).
ERR_BUDGET = 5e-7


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    if not torch.cuda.is_available():          # CPU emulation of the host logic (tests/test_dist_gloo.py): no stream
        return C.c_void_p(0)
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_SYNC_WS = {}


def attach_sync_ws(args: "_lib.X2Args", device) -> None:
    """Give a paired-SM launch its lockstep scratch (srk_x2_args.sync_ws): one 1 MB buffer per
    (device, stream) -- launches on one stream never overlap."""
    key = (str(device), torch.cuda.current_stream().cuda_stream)
    ws = _SYNC_WS.get(key)
    if ws is None:
        ws = _SYNC_WS[key] = torch.zeros(1 << 18, dtype=torch.int32, device=device)
    args.sync_ws, args.sync_ws_bytes = ws.data_ptr(), ws.numel() * 4


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("simrank_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.load()
    return torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


# --------------------------------------------------------------------------- graph build
DEVICE_CSR_MIN_EDGES = 32768          # below this the host sort beats a device round trip
_DEVICE_CSR_MAX_COLUMNS = 819200      # srk_edges_to_csr: K-bit row bitmap in shared memory


def device_csr(rows, cols, M, K, device=None):
    """Edge positions -> CSR on the GPU (srk_edges_to_csr: histogram, scan, scatter, per-row bitmap
    sort with duplicate detection): the device replacement of the pivot + row scatter of
    SimRank.py:50-52.  -> (indptr, indices, (device indptr, device indices)) or None when the graph is
    too small to be worth the round trip (the host sort takes it).  Raises the pivot's ValueError on
    a duplicate (row, column) pair."""
    from .graph import _DUPLICATE_MSG
    m = int(np.asarray(rows).size)
    if m < DEVICE_CSR_MIN_EDGES or K > _DEVICE_CSR_MAX_COLUMNS or max(M, m) >= (1 << 31) or not torch.cuda.is_available():
        return None
    dev = require_cuda(device)
    lib = _lib.load()
    r = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int32)).to(dev)
    c = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int32)).to(dev)
    indptr = torch.empty(M + 1, dtype=torch.int64, device=dev)
    indices = torch.empty(m, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = int(lib.srk_edges_to_csr_workspace(m, M))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.srk_edges_to_csr(_ptr(r), _ptr(c), m, M, K, _ptr(indptr), _ptr(indices), _ptr(status), _ptr(ws),
                                    nbytes, _stream()), "srk_edges_to_csr")
    st = int(status.item())
    if st & 2:
        raise IndexError("edge endpoint outside the node range")
    if st & 1:
        raise ValueError(_DUPLICATE_MSG)
    return indptr.cpu().numpy(), indices.cpu().numpy(), (indptr, indices)


# --------------------------------------------------------------------------- operators
class DeviceOperator:
    """``G = diag(g) * A`` resident on the GPU: CSR always, dense uint8 A on demand."""

    def __init__(self, host: HostOperator, device):
        self.host = host
        self.M, self.K = host.M, host.K
        self.device = device
        dev_csr = getattr(host, "dev_csr", None)                  # built on the device already (device_csr)
        if dev_csr is not None and dev_csr[0].device == torch.device(device):
            self.indptr, self.indices = dev_csr
        else:
            self.indptr = torch.from_numpy(host.indptr).to(device)
            self.indices = torch.from_numpy(host.indices).to(device)
        self.g_host = np.ascontiguousarray(host.g)
        self.g = torch.from_numpy(self.g_host).to(device)
        self.dead = torch.from_numpy(host.dead).to(device)
        self._a8 = None
        self._cnt16 = None

    def with_scale(self, factor: np.ndarray) -> "DeviceOperator":
        """``diag(factor) * G``: same pattern, rescaled rows (SimRank++ W = diag(spread) G,
        SimRank.py:333 -- an n^3 dgemm there, an O(n) host product here)."""
        other = object.__new__(DeviceOperator)
        other.__dict__.update(self.__dict__)
        other.g_host = np.ascontiguousarray(self.g_host * np.asarray(factor, dtype=np.float64))
        other.g = torch.from_numpy(other.g_host).to(self.device)
        return other

    @property
    def lda(self) -> int:
        return _round_up(max(self.K, 1), 128)

    def dense_u8(self) -> torch.Tensor:
        if self._a8 is None:
            a8 = torch.empty((self.M, self.lda), dtype=torch.uint8, device=self.device)
            _lib.check(_lib.load().srk_csr_to_dense_u8(_ptr(self.indptr), _ptr(self.indices), 0, self.M, self.K,
                                                       _ptr(a8), self.lda, _stream()), "srk_csr_to_dense_u8")
            self._a8 = a8
        return self._a8

    def count_bits(self) -> int:
        """16 unless two rows can have 65535 common neighbours (second-largest degree)."""
        return count_bits(self.host.deg)

    def pattern_counts(self) -> torch.Tensor:
        """Common-neighbour counts ``A A^T`` of the 0/1 pattern, the unit-diagonal term of the
        tensor-core path: ``A S A^T = A S_off A^T + A A^T``; uint16, or uint32 when a count can reach
        65535.  Held as an int16/int32 tensor (torch has no arithmetic on unsigned types; only the
        bytes matter)."""
        if self._cnt16 is None:
            bits = self.count_bits()
            ld = _round_up(max(self.M, 1), 8)
            cnt = torch.empty((self.M, ld), dtype=torch.int16 if bits == 16 else torch.int32, device=self.device)
            a8 = self.dense_u8()
            a = _lib.X2Args()
            a.mode, a.ns = _lib.SRK_X2_COUNTS, 1
            a.M, a.R, a.K = self.M, self.M, self.K
            a.A8, a.lda = a8.data_ptr(), self.lda
            a.in_planes, a.ld_in, a.in_plane_stride = a8.data_ptr(), self.lda, a8.numel()
            a.out_counts, a.ld_out_counts, a.counts_bits = cnt.data_ptr(), ld, bits
            attach_sync_ws(a, self.device)
            _lib.check(_lib.load().srk_x2_half(C.byref(a), _stream()), "srk_x2_half(COUNTS)")
            self._cnt16 = cnt
        return self._cnt16

    def row_spread(self, vals: torch.Tensor | None = None) -> torch.Tensor:
        """exp(-var) of the nonzeros of each row (SimRank.py:326-332)."""
        out = torch.empty(self.M, dtype=torch.float64, device=self.device)
        if self.M == 0:
            return out
        _lib.check(_lib.load().srk_csr_row_spread(_ptr(self.indptr), _ptr(vals), _ptr(self.g), self.M, _ptr(out),
                                                  _stream()), "srk_csr_row_spread")
        return out

    def evidence_counts(self, mode: str = "auto") -> torch.Tensor:
        """uint8 common-neighbour counts (clipped at 255) = the integer matmul of SimRank.py:315."""
        ld = _round_up(max(self.M, 1), 16)
        cnt = torch.empty((self.M, ld), dtype=torch.uint8, device=self.device)
        if self.M == 0:
            return cnt
        lib = _lib.load()
        # rows with G>0 False (g <= 0) must count as empty: only the CSR kernel knows `dead`
        has_dead = bool((self.host.dead.astype(bool) & (self.host.deg > 0)).any())
        use_i8 = not has_dead and (mode == "i8" or (mode == "auto" and lib.srk_i8_supported() and
                                                    self.M >= 2048 and self.M * self.lda <= (8 << 30)))
        if use_i8:
            a8 = self.dense_u8()
            a = _lib.X2Args()
            a.mode, a.ns = _lib.SRK_X2_COUNTS, 1
            a.M, a.R, a.K = self.M, self.M, self.K
            a.A8, a.lda = a8.data_ptr(), self.lda
            a.in_planes, a.ld_in, a.in_plane_stride = a8.data_ptr(), self.lda, a8.numel()
            a.out_counts, a.ld_out_counts, a.counts_bits = cnt.data_ptr(), ld, 8
            attach_sync_ws(a, self.device)
            _lib.check(lib.srk_x2_half(C.byref(a), _stream()), "srk_x2_half(COUNTS, uint8)")
        else:
            _lib.check(lib.srk_csr_evidence_counts(_ptr(self.indptr), _ptr(self.indices), _ptr(self.dead), self.M,
                                                   0, self.M, _ptr(cnt), ld, _stream()), "srk_csr_evidence_counts")
        return cnt


def count_bits(deg: np.ndarray) -> int:
    """Element width of the common-neighbour counts: |N(i) & N(j)| <= the second-largest degree."""
    if deg.size < 2:
        return 16
    return 16 if int(np.partition(deg, -2)[-2]) < 65535 else 32


def fixed_point_obstacle(op: HostOperator, coef: float = 0.8, lbd: float = 0.0, has_prior: bool = False):
    """Why the fixed-point (tensor-core) path cannot hold this problem, or None.  The planes store
    non-negative similarities and the kernels treat non-positive / non-finite factors as 0, so the
    path needs finite row scales g >= 0, a decay factor C > 0 and a blend weight 0 <= lbd <= 1."""
    g = np.asarray(op.g)
    if not bool(np.all(np.isfinite(g))) or not bool(np.all(g >= 0)):
        return "needs finite, non-negative 1/inNeighbors (negative or zero weight sums)"
    if not (np.isfinite(coef) and coef > 0):
        return f"needs a decay factor C > 0 (got {coef!r})"
    if has_prior and not (0.0 <= lbd <= 1.0):
        return f"needs 0 <= lbd <= 1 (got {lbd!r})"
    return None


def contraction(op: HostOperator, coef: float, blend: float = 1.0) -> float:
    """kappa = blend * C * max_i (row sum of G)^2: factor by which one update shrinks earlier errors."""
    rho = np.asarray(op.g) * op.deg
    rho_max = float(np.abs(rho).max()) if rho.size else 0.0
    return float(blend) * float(coef) * rho_max * rho_max


def choose_mode(op: HostOperator, requested: str | None = None, coef: float = 0.8, lbd: float = 0.0,
                has_prior: bool = False) -> str:
    """'csr' (exact f64 gather), 'csr16' (uint16 fixed-point gather) or 'i8' (tcgen05 fixed-point dense
    chain).  An explicit fixed-point mode that cannot hold the problem raises; 'auto' sends such
    problems -- and updates that do not contract (kappa >= 1: the a-priori error bound of
    choose_slices does not exist) -- to 'csr', and picks between the fixed-point paths by density."""
    mode = (requested or os.environ.get("SIMRANK_B200_MODE", "auto")).lower()
    if mode == "csr":
        return mode
    obstacle = fixed_point_obstacle(op, coef, lbd, has_prior)
    if mode == "csr16" and not obstacle:
        if has_prior:
            obstacle = "does not take a prior (the symmetric second half mirrors every value)"
        elif op.deg.size and int(op.deg.max()) >= (1 << 24):
            obstacle = "needs row degrees below 2^24 (exact 32-bit sums of fixed-point values)"
    if mode in ("i8", "csr16"):
        if obstacle:
            raise ValueError(f"mode={mode!r} {obstacle}; use mode='csr' (float64)")
        return mode
    if mode != "auto":
        raise ValueError(f"unknown mode {mode!r}")
    blend = (1.0 - lbd) if has_prior else 1.0
    ok = bool(_lib.load().srk_i8_supported()) and obstacle is None and contraction(op, coef, blend) < 0.999
    dense_bytes = op.M * _round_up(op.K, 128)
    if not (ok and min(op.M, op.K) >= 1024 and dense_bytes <= (16 << 30)):
        return "csr"
    # Measured crossover at BASELINE cfg4 (n = 32768, 0.19 % dense; profiles/r2_bench_n1.json): the dense
    # tensor-core chain costs 63 ms whatever the density, the fixed-point gather 26 ms, growing with
    # nnz * n -- they meet at a density of about 0.45 %.  The float64 gather is 2.3x slower than the
    # fixed-point one and only beats the dense chain below 0.1 %.
    density = op.nnz / max(1, op.M * op.K)
    if density >= 1.0 / 256:
        return "i8"
    if not has_prior and (op.deg.size == 0 or int(op.deg.max()) < (1 << 24)):
        return "csr16"
    return "i8" if density >= 1.0 / 1024 else "csr"


# --------------------------------------------------------------------------- one similarity matrix
def slice_delta(ns: int, coef: float, blend: float, rho_max: float, s_off_max: float) -> float:
    """Largest deviation ONE update adds (see choose_slices)."""
    return 1.5 * blend * coef * rho_max * rho_max * s_off_max / 256.0 ** ns


def gather_qmax(deg: np.ndarray) -> float:
    """Largest fixed-point value of the uint16 gather of an operator: the 32-bit sums over a row's
    neighbours are exact while deg * qmax < 2^32, so 65535 unless a row has more than 65536
    neighbours (then fewer levels: the same arithmetic with a coarser step)."""
    dmax = int(deg.max()) if deg.size else 0
    return float(min(65535, ((1 << 32) - 1) // max(dmax, 1)))


class ListSplit:
    """Neighbour lists of the hub rows, cut into pieces for the fixed-point gather (include/simrank_b200.h,
    "Split neighbour lists").  A ratings graph has rows with tens of thousands of neighbours next to rows
    with a handful (BASELINE cfg5: the popular items): one warp walking such a row is the tail of the whole
    launch.  Rows with at least ``min_deg`` neighbours are cut (a) at the boundaries of ``ranges`` equal
    ranges of X rows -- pieces that run at the same time then gather from one range, which stays in L2 when
    the whole panel (rows of X x 1 KB) does not -- and (b) into pieces of at most ``piece`` neighbours.
    ``srk_csr_half`` SRK_CSR_ACCUM sums the pieces into ``accum[slot]`` (exact integers: the order does not
    matter, results are bit-identical to the unsplit call) and the FIRST / FINAL launch starts those rows
    from the sums with an empty list.

    Tunables (environment, read when the plan is made): SRK_SPLIT_MIN (0 = never split), SRK_SPLIT_PIECE,
    SRK_SPLIT_RANGE_MB (32: a range of X rows is at most this many MB of 1 KB segments).  Defaults from
    profiles/r2_csr_split_shapes_real.jsonl (one rank's launches of BASELINE cfg5 replayed on one GPU):
    rows of 1024+ neighbours in pieces of 256 when a panel of X does not fit in L2 (S2's halves: 201 -> 60
    ms and 41 -> 12 ms), rows of 256+ in pieces of 512 when it does (S1's halves: 91 -> 80 and 13.3 -> 11.8
    ms; there the gain is balance inside the CTAs, whose rows differ tenfold in length)."""

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, K: int, min_deg: int, piece: int, ranges: int,
                 all_rows: bool = False):
        dev = indptr.device
        M = indptr.numel() - 1
        deg = indptr[1:] - indptr[:-1]
        # all_rows: EVERY list goes through the pieces and slot = row (the sums of the whole operator end up in
        # accum, for SRK_CSR_FINISH); otherwise only the hub rows, numbered in order
        hub = deg >= (1 if all_rows else min_deg)
        self.all_rows = bool(all_rows)
        self.rows = M if all_rows else int(hub.sum().item())
        self.row_lo = indptr[:-1].clone()
        self.row_hi = torch.where(hub, self.row_lo, indptr[1:])                      # hubs: nothing left to gather
        if all_rows:
            slot = torch.arange(M, dtype=torch.int32, device=dev)
        else:
            slot = torch.cumsum(hub.to(torch.int32), 0, dtype=torch.int32) - 1
        self.slot = torch.where(hub, slot, torch.full_like(slot, -1))
        hub_rows = torch.nonzero(hub).flatten()
        hub_deg = deg[hub_rows]
        # positions (in `indices`) of the hub rows' entries, row by row
        starts = indptr[:-1][hub_rows]
        first = torch.cumsum(hub_deg, 0) - hub_deg
        owner = torch.repeat_interleave(torch.arange(hub_rows.numel(), device=dev), hub_deg)
        pos = starts[owner] + (torch.arange(int(hub_deg.sum().item()), device=dev) - first[owner])
        span = max(1, -(-max(K, 1) // max(ranges, 1)))
        key = owner * ranges + torch.div(indices[pos].to(torch.int64), span, rounding_mode="floor")
        head = torch.ones_like(key, dtype=torch.bool)
        head[1:] = key[1:] != key[:-1]                                              # first entry of a (row, range) segment
        seg_at = torch.nonzero(head).flatten()
        seg_lo = pos[seg_at]
        seg_len = torch.diff(seg_at, append=torch.tensor([key.numel()], device=dev))
        seg_key = key[seg_at]
        n_piece = torch.div(seg_len + piece - 1, piece, rounding_mode="floor")
        size = torch.div(seg_len + n_piece - 1, n_piece, rounding_mode="floor")     # equal pieces, no short tail
        pseg = torch.repeat_interleave(torch.arange(seg_at.numel(), device=dev), n_piece)
        pfirst = torch.cumsum(n_piece, 0) - n_piece
        k = torch.arange(pseg.numel(), device=dev) - pfirst[pseg]
        lo = seg_lo[pseg] + k * size[pseg]
        hi = torch.minimum(lo + size[pseg], (seg_lo + seg_len)[pseg])
        powner = torch.div(seg_key[pseg], ranges, rounding_mode="floor")            # hub row (position in hub_rows) of a piece
        pslot = slot[hub_rows][powner].to(torch.int64)
        per_row = torch.zeros(M, dtype=torch.int64, device=dev)
        per_row.index_add_(0, hub_rows[powner], torch.ones_like(powner))
        alone = per_row[hub_rows[powner]] == 1                                      # the piece is its row's whole list
        pslot = torch.where(alone, -pslot - 1, pslot)                               # stored, not added
        # range by range (pieces that run together gather from one range of X); inside a range the long pieces
        # first, so that the 8 pieces of a CTA are alike and the short ones fill the end of the launch
        order = torch.argsort((seg_key[pseg] % ranges) * (1 << 32) - (hi - lo), stable=True)
        self.piece_lo, self.piece_hi = lo[order].contiguous(), hi[order].contiguous()
        self.piece_slot = pslot[order].to(torch.int32).contiguous()
        self.pieces = int(self.piece_lo.numel())
        # slots that are added to (or never written: empty lists) start from zero
        needs_zero = per_row != 1
        self.zero_slots = slot[needs_zero & (hub | all_rows)].to(torch.int64) if all_rows else None
        self.ranges, self.piece, self.min_deg = ranges, piece, min_deg
        self._accum = None

    @classmethod
    def plan(cls, indptr: torch.Tensor, indices: torch.Tensor, K: int, max_deg: int, all_rows: bool = False):
        """A split for this operator, or None when no row is long enough to need one (``all_rows``: always a
        plan, every list in pieces)."""
        fits = K * 1024 <= 64 * 2 ** 20                               # a 1 KB-wide panel of X stays in L2
        min_deg = int(os.environ.get("SRK_SPLIT_MIN", "256" if fits else "1024"))
        if not all_rows and (min_deg <= 0 or max_deg < min_deg):
            return None
        piece = max(4, int(os.environ.get("SRK_SPLIT_PIECE", "512" if fits else "256")))
        range_mb = float(os.environ.get("SRK_SPLIT_RANGE_MB", "32"))
        ranges = max(1, int(np.ceil(K * 1024 / (range_mb * 2 ** 20))))
        return cls(indptr, indices, K, min_deg, piece, ranges, all_rows)

    def accumulate(self, lib, indices_ptr, x_ptr, ldx: int, L: int, K: int, qmax: float, upper: bool = False) -> None:
        """Zero the sums and add every piece's column sums of X[:, :L] (one launch).  ``upper`` (all_rows plans
        of a square operator): a piece of row i skips the column panels entirely left of column i -- the
        symmetric second half only needs the pairs r >= i."""
        ld = _round_up(max(L, 1), 512)
        if self._accum is None or self._accum.shape[1] < ld:
            self._accum = torch.empty((self.rows, ld), dtype=torch.int32, device=self.piece_lo.device)
        if self.zero_slots is None:
            self._accum.zero_()
        elif self.zero_slots.numel():
            self._accum.index_fill_(0, self.zero_slots, 0)
        a = _lib.CsrArgs()
        a.elem, a.mode = _lib.SRK_ELEM_U16, _lib.SRK_CSR_ACCUM
        a.indices = indices_ptr
        a.M, a.row_begin, a.row_end = self.pieces, 0, self.pieces
        a.X, a.ldx, a.L, a.K = x_ptr, ldx, L, K
        a.qmax = qmax
        a.symmetric = 1 if upper else 0
        self.attach(a, pieces=True)
        _lib.check(lib.srk_csr_half(C.byref(a), _stream()), "srk_csr_half(u16, accum)")

    def attach(self, a: "_lib.CsrArgs", pieces: bool = False) -> None:
        if pieces:
            a.row_lo, a.row_hi, a.accum_slot = self.piece_lo.data_ptr(), self.piece_hi.data_ptr(), self.piece_slot.data_ptr()
        else:
            a.row_lo, a.row_hi, a.accum_slot = self.row_lo.data_ptr(), self.row_hi.data_ptr(), self.slot.data_ptr()
        a.accum, a.ld_accum = self._accum.data_ptr(), self._accum.shape[1]


def choose_slices(requested, coef: float, blend: float, rho_max: float, s_off_max: float) -> int:
    """Planes per matrix for one update of the tensor-core path.

    Rounding happens twice per update, when S_off and U = A S_off are cut into NS uint8 planes with
    bounds b_S(r) = max_k S_off[r,k] (exact) and b_U(j) = deg_j * max(S_off) rounded up to a power
    of two (< 2 b_U): at most half a step of bound / 256^NS each, i.e. 0.5 and < 1 steps of the
    tight bounds.  Propagated through ``coef * g g^T o (A . A^T)`` the two add up to at most
    ``delta = 1.5 * blend * coef * rho_max^2 * max(S_off) / 256^NS`` per update (rho = row sums of
    G, 1 for unweighted graphs), and the update contracts earlier errors by kappa = blend * coef *
    rho_max^2, so the deviation from the float64 iteration never exceeds delta / (1 - kappa).  The
    smallest NS in {2, 3, 4} that keeps this below ERR_BUDGET is used; an integer request is
    honoured as is.  When the update does not contract (kappa >= 1; 'auto' mode sends those graphs
    to the float64 path, engine.choose_mode) there is no such series: 4 planes are used and the bound
    that holds after k updates is the recursion the solver tracks (``_Half.err``, FitInfo.error_bound)."""
    if requested not in (None, "auto"):
        ns = int(requested)
        if ns not in (2, 3, 4):
            raise ValueError("slices must be 2, 3, 4 or 'auto'")
        return ns
    kappa = blend * coef * rho_max * rho_max
    if not kappa < 0.999:
        return 4
    for ns in (2, 3, 4):
        if slice_delta(ns, coef, blend, rho_max, s_off_max) / (1.0 - kappa) <= ERR_BUDGET:
            return ns
    return 4


class _Half:
    """State for updating ONE similarity matrix S_out (n_out x n_out) from S_in (n_in x n_in)
    through ``op`` (n_out x n_in):  S_out <- epilogue(coef * G S_in G^T).  The directed classes
    use one instance with S_in is S_out; the bipartite classes use two (SimRank.py:297-302)."""

    def __init__(self, op: DeviceOperator, coef: float, mode: str, ns, evidence=None, prior=None, lbd=0.0,
                 evidence_from_pattern=False):
        self.op, self.coef, self.mode, self.ns = op, float(coef), mode, ns
        self.n_out, self.n_in = op.M, op.K
        dev = op.device
        self.ld = _round_up(max(self.n_out, 1), 16)
        self.S = torch.empty((self.n_out, self.ld), dtype=torch.float64, device=dev)
        lib = _lib.load()
        if self.n_out:                                            # an empty graph has an empty S (null data pointer)
            _lib.check(lib.srk_set_identity_f64(_ptr(self.S), self.ld, self.n_out, self.n_out, 0, _stream()),
                       "srk_set_identity_f64")
        self.evidence, self.prior, self.lbd = evidence, prior, float(lbd)
        self.scal = torch.zeros(2, dtype=torch.float64, device=dev)        # [maxdiff, maxoff]
        self.events = None          # set to a list to collect (name, start, end) CUDA events per launch
        self.maxoff = 0.0                                                  # max off-diagonal of current S
        self.slices_used = []                                              # NS of every update (i8 mode)
        self.err = 0.0              # guaranteed max-abs deviation of S from the float64 iteration (i8 mode)
        if mode == "csr":
            self.T = None                                                  # allocated by the first update
            return
        host = op.host
        if mode == "csr16":
            # uint16 matrices are held as int16 tensors (only the bytes matter)
            self.rho = op.g_host * host.deg
            self.rho_max = float(self.rho.max()) if self.rho.size else 0.0
            self.ldxt = _round_up(max(self.n_out, 1), 64)                  # Xq of the own S (as a source)
            self.Xq, self.unit = None, torch.zeros(max(self.n_out, 1), dtype=torch.float64, device=dev)
            self.ldt = _round_up(max(self.n_out, 1), 64)
            self.Tq = torch.empty((self.n_in, self.ldt), dtype=torch.int16, device=dev)
            self.deg_dev = torch.from_numpy(host.deg.astype(np.float64)).to(dev)
            self.qmax = gather_qmax(host.deg)                              # of the gathers THIS half runs
            self.evidence_from_pattern = bool(evidence_from_pattern)
            self.counts = op.pattern_counts()
            self.version, self._quantized_version = 0, (-1, 0.0)
            # hub rows (tens of thousands of neighbours) are pre-summed in pieces, see ListSplit
            self.split = ListSplit.plan(op.indptr, op.indices, op.K, int(host.deg.max()) if host.deg.size else 0)
            # Both halves as SRK_CSR_ACCUM over EVERY list + a streaming pass (FINISH_FIRST; symmetric FINISH
            # over the pairs r >= i): the gather launch carries no tile and no epilogue and runs at the L2
            # roof (DESIGN.md "K3").  SRK_CSR_VIA_ACCUM=0 keeps the fused launches.
            self.split_all = None
            if os.environ.get("SRK_CSR_VIA_ACCUM", "1") == "1" and host.nnz:
                self.split_all = ListSplit.plan(op.indptr, op.indices, op.K, int(host.deg.max()), all_rows=True)
            return
        self.rho = op.g_host * host.deg                                    # row sums of G
        self.rho_max = float(self.rho.max()) if self.rho.size else 0.0
        self.prior_max = float(prior.max()) if prior is not None else 0.0
        self.ldp = _round_up(max(self.n_out, 1), 128)                      # planes of S_out
        self.ldu = _round_up(max(self.n_in, 1), 128)                       # planes of U (n_out x n_in)
        self.a8 = op.dense_u8()
        self.deg_dev = torch.from_numpy(host.deg.astype(np.float64)).to(dev)
        # ---- paired-SM path: planes are cut from S right before they are used, with exact bounds
        self.ns_alloc = 3 if ns in (None, "auto") else int(ns)
        self.planes = None                                                 # allocated on first use as a source
        self.planes_U = torch.empty((self.ns_alloc, self.n_out, self.ldu), dtype=torch.uint8, device=dev)
        self.bound_vec = torch.zeros(max(self.n_out, 1), dtype=torch.float64, device=dev)
        # keys of the row maxima of the current S_off, collected by the FINAL epilogue (S = I: all zero)
        self.rowmax_hi = torch.zeros(max(self.n_out, 1), dtype=torch.int32, device=dev)
        self.evidence_from_pattern = bool(evidence_from_pattern)
        self.counts = op.pattern_counts()                                  # uint16 A A^T, [n_out, ldc]
        self.version = 0                                                   # bumped by every update of S
        self._sliced = (-1, 0)                                             # (version, ns) of self.planes

    # -- epilogue description shared by both modes
    def _epilogue(self) -> _lib.Epilogue:
        e = _lib.Epilogue()
        e.coef = self.coef
        if self.evidence is not None and not getattr(self, "evidence_from_pattern", False):
            e.evidence, e.ld_evidence = self.evidence.data_ptr(), self.evidence.stride(0)
        if self.prior is not None:
            e.prior, e.ld_prior, e.lambda_ = self.prior.data_ptr(), self.prior.stride(0), self.lbd
        e.s_old, e.ld_s_old = self.S.data_ptr(), self.ld
        e.maxdiff = self.scal.data_ptr()
        e.maxoff = self.scal.data_ptr() + 8
        return e

    def _timed(self, name, launch):
        """Run ``launch()``; when profiling, bracket it with CUDA events on the launching stream."""
        if self.events is None:
            return launch()
        st = torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        rc = launch()
        b.record(st)
        self.events.append((name, a, b))
        return rc

    def _planes_for(self, ns: int) -> torch.Tensor:
        """Planes of the off-diagonal part of the CURRENT S, bounded per row by the maxima the FINAL
        epilogue collected while writing it (srk_slice_rows_key_f64: one pass); cached per
        (version of S, ns)."""
        if self.planes is None or self.planes.shape[0] < ns:
            self.planes = torch.empty((max(ns, self.ns_alloc), self.n_out, self.ldp), dtype=torch.uint8,
                                      device=self.S.device)
            self._sliced = (-1, 0)
        if self._sliced != (self.version, ns):
            lib = _lib.load()
            _lib.check(self._timed("slice_rows_key", lambda: lib.srk_slice_rows_key_f64(
                _ptr(self.S), self.ld, self.n_out, self.n_out, 0, ns, _ptr(self.rowmax_hi), _ptr(self.planes),
                self.ldp, self.planes.stride(0), _ptr(self.bound_vec), _stream())), "srk_slice_rows_key_f64")
            self._sliced = (self.version, ns)
        return self.planes

    def update(self, src: "_Half") -> None:
        """Launch the two half-products (asynchronous).  ``src`` holds S_in (its S / planes)."""
        lib = _lib.load()
        self.scal.zero_()
        if self.mode == "csr":
            return self._update_f64(src)
        if self.mode == "csr16":
            return self._update_csr16(src)
        # ---- paired-SM tensor-core path
        blend = (1.0 - self.lbd) if self.prior is not None else 1.0
        ns = choose_slices(self.ns, self.coef, blend, self.rho_max, src.maxoff)
        if ns > self.planes_U.shape[0]:
            self.planes_U = torch.empty((ns, self.n_out, self.ldu), dtype=torch.uint8, device=self.S.device)
        self.slices_used.append(ns)
        # e_new <= kappa * e(S_in) + delta: the update maps a perturbation of S_in of size e to at most
        # kappa * e, and its own two roundings add at most delta (choose_slices)
        self._err_next = blend * self.coef * self.rho_max ** 2 * src.err + \
            slice_delta(ns, self.coef, blend, self.rho_max, src.maxoff)
        planes_in = src._planes_for(ns)
        # U[j, r] = sum_{k in N(j)} S_off[r, k] <= deg_j * max(S_off); the guard keeps a value that
        # attains the bound below the last level (>= 256^NS / (256^NS - 1) for every NS >= 2)
        guard = 1.0 + 2.0 ** -14
        bound_U = _lib.RowBound.of(self.deg_dev.data_ptr(), src.maxoff * guard, 0.0)

        a = _lib.X2Args()
        a.mode, a.ns = _lib.SRK_X2_MID, ns
        a.M, a.R, a.K = self.n_out, self.n_in, self.n_in
        a.A8, a.lda = self.a8.data_ptr(), self.op.lda
        a.in_planes, a.ld_in, a.in_plane_stride = planes_in.data_ptr(), src.ldp, planes_in.stride(0)
        a.in_rowbound = _lib.RowBound.of(src.bound_vec.data_ptr(), 1.0, 0.0)
        a.out_planes, a.ld_outp, a.out_plane_stride = self.planes_U.data_ptr(), self.ldu, self.planes_U.stride(0)
        a.out_rowbound = bound_U
        attach_sync_ws(a, self.S.device)
        _lib.check(self._timed("x2_half_mid", lambda: lib.srk_x2_half(C.byref(a), _stream())), "srk_x2_half(MID)")

        b = _lib.X2Args()
        b.mode, b.ns = _lib.SRK_X2_FINAL, ns
        # the mirrored store needs a symmetric epilogue: evidence counts are, an arbitrary prior is not
        b.layout = _lib.SRK_X2_SYMMETRIC if self.prior is None else _lib.SRK_X2_DIRECT
        b.M, b.R, b.K = self.n_out, self.n_out, self.n_in
        b.A8, b.lda = self.a8.data_ptr(), self.op.lda
        b.in_planes, b.ld_in, b.in_plane_stride = self.planes_U.data_ptr(), self.ldu, self.planes_U.stride(0)
        b.in_rowbound = bound_U
        b.g_a = b.g_v = self.op.g.data_ptr()
        b.counts, b.ld_counts, b.add_counts = self.counts.data_ptr(), self.counts.stride(0), 1
        b.counts_bits = 8 * self.counts.element_size()
        b.use_evidence = 1 if self.evidence_from_pattern else 0
        b.out_f64, b.ld_out, b.diag_offset = self.S.data_ptr(), self.ld, 0
        self.rowmax_hi.zero_()                                     # after the slice above read the old keys
        b.rowmax_hi = self.rowmax_hi.data_ptr()
        b.epi = self._epilogue()
        attach_sync_ws(b, self.S.device)
        _lib.check(self._timed("x2_half_final", lambda: lib.srk_x2_half(C.byref(b), _stream())),
                   "srk_x2_half(FINAL)")
        self.version += 1

    def _update_f64(self, src: "_Half") -> None:
        """Float64 CSR update: T = (G S_in)^T, then S = epilogue((G T)^T)."""
        lib, op = _lib.load(), self.op
        if getattr(self, "T", None) is None:
            self.ldt64 = _round_up(max(self.n_out, 1), 16)
            self.T = torch.empty((self.n_in, self.ldt64), dtype=torch.float64, device=self.S.device)
        a = _lib.CsrArgs()
        a.elem, a.mode = _lib.SRK_ELEM_F64, _lib.SRK_CSR_FIRST
        a.indptr, a.indices, a.g = op.indptr.data_ptr(), op.indices.data_ptr(), op.g.data_ptr()
        a.M, a.row_begin, a.row_end = op.M, 0, op.M
        a.X, a.ldx, a.L, a.K = src.S.data_ptr(), src.ld, self.n_in, self.n_in
        a.OUT, a.ldo = self.T.data_ptr(), self.ldt64
        _lib.check(self._timed("csr_half_first", lambda: lib.srk_csr_half(C.byref(a), _stream())),
                   "srk_csr_half(f64, first)")
        b = _lib.CsrArgs()
        b.elem, b.mode = _lib.SRK_ELEM_F64, _lib.SRK_CSR_FINAL
        # without a prior the result is symmetric: each unordered pair is computed once and mirrored
        b.symmetric = 1 if self.prior is None and os.environ.get("SIMRANK_B200_CSR_SYMMETRIC", "1") != "0" else 0
        b.indptr, b.indices, b.g = op.indptr.data_ptr(), op.indices.data_ptr(), op.g.data_ptr()
        b.M, b.row_begin, b.row_end = op.M, 0, op.M
        b.X, b.ldx, b.L, b.K = self.T.data_ptr(), self.ldt64, self.n_out, self.n_in
        b.OUT, b.ldo = self.S.data_ptr(), self.ld
        if getattr(self, "evidence_from_pattern", False):          # a float64 update of the csr16 mode
            b.counts, b.ld_counts = self.counts.data_ptr(), self.counts.stride(0)
            b.counts_bits, b.use_evidence = 8 * self.counts.element_size(), 1
        b.epi = self._epilogue()
        _lib.check(self._timed("csr_half_final", lambda: lib.srk_csr_half(C.byref(b), _stream())),
                   "srk_csr_half(f64, second)")
        self.version = getattr(self, "version", 0) + 1

    def _quantized(self, qmax: float = 65535.0):
        """uint16 source operand of the CURRENT S for the fixed-point gather (cached per version of S and
        range): Xq[k, r] = rint(S_off[r, k] / unit[r]), unit[r] = row maximum / qmax; ``qmax`` is the
        consumer's (gather_qmax of the operator whose rows sum these values)."""
        if self.Xq is None:
            self.Xq = torch.empty((self.n_out, self.ldxt), dtype=torch.int16, device=self.S.device)
        if self._quantized_version != (self.version, qmax):
            lib = _lib.load()
            _lib.check(self._timed("quantize_rows_u16", lambda: lib.srk_quantize_rows_u16(
                _ptr(self.S), self.ld, self.n_out, self.n_out, 0, _ptr(self.Xq), self.ldxt, _ptr(self.unit),
                qmax, 1, _stream())), "srk_quantize_rows_u16")      # S is bit-exactly symmetric in this mode
            self._quantized_version = (self.version, qmax)
        return self.Xq, self.unit

    def _update_csr16(self, src: "_Half") -> None:
        """Fixed-point CSR path: the two gathers sum uint16 values as exact integers."""
        lib, op = _lib.load(), self.op
        # Same two roundings as the tensor-core path with 2 planes (bounds: exact row maximum of S_off,
        # deg * max(S_off) for U -- not even rounded up to a power of two here).  When 16 bits cannot
        # keep the guaranteed deviation inside ERR_BUDGET (large similarities, C close to 1) THIS update
        # runs in float64: both kinds of update read and write the same float64 S.
        kappa = self.coef * self.rho_max ** 2
        # (a coarser range -- qmax < 65535 -- scales the 16-bit bound accordingly)
        if getattr(self, "force_f64", False) or choose_slices(self.ns, self.coef, 1.0, self.rho_max,
                                                              src.maxoff * 65536.0 / (self.qmax + 1.0)) > 2:
            self.slices_used.append(0)                                     # 0 = float64 update
            self._err_next = kappa * src.err
            return self._update_f64(src)                                   # bumps self.version
        self.slices_used.append(2)
        qmax = self.qmax
        self._err_next = kappa * src.err + slice_delta(2, self.coef, 1.0, self.rho_max, src.maxoff) * 65536.0 / (qmax + 1.0)
        xq, unit = src._quantized(qmax)
        guard = 1.0 + 2.0 ** -14
        a = _lib.CsrArgs()
        a.elem, a.mode = _lib.SRK_ELEM_U16, _lib.SRK_CSR_FIRST
        a.indptr, a.indices, a.g = op.indptr.data_ptr(), op.indices.data_ptr(), op.g.data_ptr()
        a.M, a.row_begin, a.row_end = op.M, 0, op.M
        a.X, a.ldx, a.L, a.K = xq.data_ptr(), src.ldxt, self.n_in, self.n_in
        a.OUT, a.ldo = self.Tq.data_ptr(), self.ldt
        a.in_unit = _lib.RowBound.of(unit.data_ptr(), 1.0, 0.0)
        a.out_bound = _lib.RowBound.of(self.deg_dev.data_ptr(), src.maxoff * guard, 0.0)
        a.qmax = qmax
        via = self.split_all

        def first_via_accum():
            self._timed("csr16_first_accum", lambda: via.accumulate(lib, a.indices, a.X, a.ldx, a.L, a.K, qmax))
            a.mode, a.accum, a.ld_accum = _lib.SRK_CSR_FINISH_FIRST, via._accum.data_ptr(), via._accum.shape[1]
            return self._timed("csr16_first_finish", lambda: lib.srk_csr_half(C.byref(a), _stream()))
        if via is not None:
            _lib.check(self._timed("csr16_half_first", first_via_accum), "srk_csr_half(u16, first via accum)")
        else:
            if self.split is not None:
                self._timed("csr16_accum", lambda: self.split.accumulate(lib, a.indices, a.X, a.ldx, a.L, a.K, qmax))
                self.split.attach(a)
            _lib.check(self._timed("csr16_half_first", lambda: lib.srk_csr_half(C.byref(a), _stream())),
                       "srk_csr_half(u16, first)")
        b = _lib.CsrArgs()
        b.elem, b.mode, b.symmetric = _lib.SRK_ELEM_U16, _lib.SRK_CSR_FINAL, 1
        b.indptr, b.indices, b.g = a.indptr, a.indices, a.g
        b.M, b.row_begin, b.row_end = op.M, 0, op.M
        b.X, b.ldx, b.L, b.K = self.Tq.data_ptr(), self.ldt, self.n_out, self.n_in
        b.OUT, b.ldo = self.S.data_ptr(), self.ld
        b.in_unit = _lib.RowBound.of(self.deg_dev.data_ptr(), src.maxoff * guard / qmax, 0.0)
        b.qmax = qmax
        b.g_col = op.g.data_ptr()
        b.counts, b.ld_counts = self.counts.data_ptr(), self.counts.stride(0)
        b.counts_bits, b.add_counts = 8 * self.counts.element_size(), 1
        b.use_evidence = 1 if self.evidence_from_pattern else 0
        b.epi = self._epilogue()

        def second_via_accum():
            self._timed("csr16_second_accum",
                        lambda: via.accumulate(lib, b.indices, b.X, b.ldx, b.L, b.K, qmax, upper=True))
            b.mode, b.accum, b.ld_accum = _lib.SRK_CSR_FINISH, via._accum.data_ptr(), via._accum.shape[1]
            return self._timed("csr16_second_finish", lambda: lib.srk_csr_half(C.byref(b), _stream()))
        if via is not None:
            _lib.check(self._timed("csr16_half_final", second_via_accum), "srk_csr_half(u16, second via accum)")
            self.version += 1
            return
        if self.split is not None:
            self._timed("csr16_accum", lambda: self.split.accumulate(lib, b.indices, b.X, b.ldx, b.L, b.K, qmax))
            self.split.attach(b)
        _lib.check(self._timed("csr16_half_final", lambda: lib.srk_csr_half(C.byref(b), _stream())),
                   "srk_csr_half(u16, second)")
        self.version += 1

    def finish(self) -> float:
        """Read back max|dS| (host sync) and commit the range of the new S."""
        maxdiff, maxoff = self.scal.tolist()
        self.maxoff = maxoff
        if self.mode != "csr":
            self.err = self._err_next
        return maxdiff

    def result(self) -> torch.Tensor:
        return self.S[:, : self.n_out]

    # the names the row-sharded halves use (dist.py), so that callers can treat both alike
    local_result = result
    row0 = 0

    @property
    def rows(self) -> int:
        return self.n_out


@dataclass
class FitInfo:
    applied: int          # updates performed (== k of "Converged at iteration k")
    converged: bool
    last_maxdiff: tuple
    mode: str
    slices_used: tuple = ()      # i8 mode: planes of every update, per matrix
    error_bound: tuple = ()      # i8 mode: guaranteed max-abs deviation from the float64 iteration, per matrix


class DirectedSolver:
    """``S <- [E o] C * W S W^T [blended with a prior]; diag <- 1`` (SimRank.py:139, :361, :453)."""

    def __init__(self, op: DeviceOperator, C_: float, evidence=None, prior=None, lbd=0.0, mode=None, ns=_NS_DEFAULT,
                 evidence_from_pattern=False):
        self.mode = choose_mode(op.host, mode, C_, lbd, prior is not None)
        self.half = _Half(op, C_, self.mode, ns, evidence, prior, lbd, evidence_from_pattern)

    def step(self) -> float:
        self.half.update(self.half)
        return self.half.finish()

    @property
    def S(self) -> torch.Tensor:
        return self.half.result()


class BipartiteSolver:
    """Gauss-Seidel alternation of SimRank.py:297-302 (and :419-424, :487-492)."""

    def __init__(self, op12: DeviceOperator, op21: DeviceOperator, C1: float, C2: float, evidence1=None,
                 evidence2=None, prior1=None, prior2=None, lbd1=0.0, lbd2=0.0, mode=None, ns=_NS_DEFAULT,
                 evidence1_from_pattern=False, evidence2_from_pattern=False):
        m1 = choose_mode(op12.host, mode, C1, lbd1, prior1 is not None)
        m2 = choose_mode(op21.host, mode, C2, lbd2, prior2 is not None)
        self.mode = m1 if m1 == m2 else "csr"
        self.h1 = _Half(op12, C1, self.mode, ns, evidence1, prior1, lbd1, evidence1_from_pattern)   # S1 from S2 (G12)
        self.h2 = _Half(op21, C2, self.mode, ns, evidence2, prior2, lbd2, evidence2_from_pattern)   # S2 from S1 (G21)

    def step(self):
        self.h1.update(self.h2)
        d1 = self.h1.finish()
        self.h2.update(self.h1)            # uses the NEW S1 (SimRank.py:301)
        d2 = self.h2.finish()
        return d1, d2

    @property
    def S1(self) -> torch.Tensor:
        return self.h1.result()

    @property
    def S2(self) -> torch.Tensor:
        return self.h2.result()


def run_loop(step, iterations: int, eps: float, pair: bool, on_iteration=None, initial=None):
    """The reference's loop skeleton (SimRank.py:129-140 / 288-302): test convergence BEFORE each
    update, using max|dS| of the previous update (``|I - 0|`` = 1 before the first; ``initial``
    overrides that: an empty matrix has no entry that differs, so an empty graph is converged at
    iteration 0 as in the reference)."""
    last = tuple(initial) if initial is not None else ((1.0, 1.0) if pair else (1.0,))
    applied, conv = 0, False
    for it in range(iterations):
        if all(not (d > eps) for d in last):
            conv = True
            break
        if on_iteration is not None:
            on_iteration(it)
        out = step()
        last = tuple(out) if pair else (out,)
        applied += 1
    return applied, conv, last


def topk_rows(S: torch.Tensor, k: int):
    """Row-wise top-k (stable, ties -> lower column).  -> (idx int32 [R,k], vals f64 [R,k])."""
    R, n = S.shape
    idx = torch.empty((R, k), dtype=torch.int32, device=S.device)
    vals = torch.empty((R, k), dtype=torch.float64, device=S.device)
    _lib.check(_lib.load().srk_topk_rows(_ptr(S), S.stride(0), R, n, k, _ptr(idx), _ptr(vals), _stream()),
               "srk_topk_rows")
    return idx, vals
