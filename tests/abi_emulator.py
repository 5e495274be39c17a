"""Host-memory emulator of ``srk_x2_half``, ``srk_slice_rows_max_f64``, ``srk_csr_half_f64``, ``srk_csr_half``
(the calls of the row-sharded fixed-point CSR solver) and ``srk_quantize_rows_u16`` (include/simrank_b200.h)
in numpy.  TEST ONLY.

It interprets the very structs the product passes to the CUDA library, but on CPU tensors, so the
multi-rank host logic (shard offsets, block assignment of the symmetric update, staging layouts,
bounds) can run under gloo without a GPU.  It is never imported by the package."""
import ctypes as C

import numpy as np

from simrank_b200 import _lib


def _view(ptr, ctype, rows, cols, ld):
    """2-D numpy view [rows, cols] of a row-major matrix with leading dimension ld (elements)."""
    n = (rows - 1) * ld + cols
    raw = np.ctypeslib.as_array((ctype * n).from_address(ptr))
    return np.lib.stride_tricks.as_strided(raw, (rows, cols), (ld * raw.itemsize, raw.itemsize))


def _f64(ptr, n):
    return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))


def _bound(rb, n):
    if rb.vec:
        return _f64(rb.vec, n) * rb.mul + rb.add
    return np.full(n, rb.add)


def pow2_exponent(b, ns):
    """f with 2^f the smallest power of two >= b, clamped to f >= 8 ns - 46; None-like -2^30 for b <= 0."""
    b = np.asarray(b, dtype=np.float64)
    m, e = np.frexp(np.where(b > 0, b, 1.0))
    f = np.maximum(np.where(m == 0.5, e - 1, e), 8 * ns - 40)
    return np.where(b > 0, f, -(1 << 30)).astype(np.int64)


def _read_planes(a):
    q = np.zeros((a.R, a.K), dtype=np.int64)
    for s in range(a.ns):
        if a.in_kblock > 0:
            kb = a.in_kblock
            for b in range(a.K // kb):
                base = a.in_planes + b * a.in_kblock_stride + s * a.in_plane_stride
                q[:, b * kb:(b + 1) * kb] += _view(base, C.c_uint8, a.R, kb, a.ld_in).astype(np.int64) << (8 * (a.ns - 1 - s))
        else:
            q += _view(a.in_planes + s * a.in_plane_stride, C.c_uint8, a.R, a.K, a.ld_in).astype(np.int64) << (8 * (a.ns - 1 - s))
    return q


def srk_x2_half(a: _lib.X2Args) -> None:
    ns, M, R, K = a.ns, a.M, a.R, a.K
    A = _view(a.A8, C.c_uint8, M, K, a.lda).astype(np.int64)
    D = A @ _read_planes(a).T                                     # [M, R] exact integers
    if a.mode == _lib.SRK_X2_COUNTS:
        if a.counts_bits == 32:
            _view(a.out_counts, C.c_uint32, M, R, a.ld_out_counts)[:, :] = D.astype(np.uint32)
        else:
            _view(a.out_counts, C.c_uint16, M, R, a.ld_out_counts)[:, :] = np.minimum(D, 65535).astype(np.uint16)
        return
    qmax = 256 ** ns
    inb = _bound(a.in_rowbound, R)
    if a.mode == _lib.SRK_X2_MID:
        fj = pow2_exponent(_bound(a.out_rowbound, M), ns)
        live = fj > -(1 << 29)
        scale = np.where(live, 2.0 ** -np.where(live, fj, 0).astype(np.float64), 0.0)
        q = np.clip(np.rint((D * scale[:, None]) * np.maximum(inb, 0.0)[None, :]), 0, qmax - 1).astype(np.int64)
        cols = -(-R // 16) * 16                                   # whole 16-column chunks are written
        for s in range(ns):
            out = _view(a.out_planes + s * a.out_plane_stride, C.c_uint8, M, cols, a.ld_outp)
            out[:, :R] = ((q >> (8 * (ns - 1 - s))) & 0xFF).astype(np.uint8)
            out[:, R:] = 0
        return
    # ---- FINAL, in the (j, r) frame of the kernel
    e = a.epi
    trans = a.layout == _lib.SRK_X2_TRANSPOSED
    sym = a.layout == _lib.SRK_X2_SYMMETRIC
    rows, cols = (R, M) if trans else (M, R)

    def mat(ptr, ctype, ld):                                      # caller's matrix in the (j, r) frame
        v = _view(ptr, ctype, rows, cols, ld)
        return v.T if trans else v

    fr = pow2_exponent(inb, ns)
    s = np.where(fr > -(1 << 29), 8 * ns - fr, 0)
    dl = np.minimum(np.maximum(-s, 0), 17)
    s = np.maximum(s, 0)
    T = D << dl[None, :]
    ctype = C.c_uint32 if a.counts_bits == 32 else C.c_uint16
    cnt = mat(a.counts, ctype, a.ld_counts).astype(np.int64) if a.counts else np.zeros((M, R), dtype=np.int64)
    if a.add_counts:
        T = T + (cnt << s[None, :])
    g_a, g_v = _f64(a.g_a, M), _f64(a.g_v, R)
    val = (T.astype(np.float64) * ((e.coef * g_v) * 2.0 ** -s.astype(np.float64))[None, :]) * g_a[:, None]
    if a.use_evidence:
        val = val * (1 - 0.5 ** np.minimum(cnt, 60).astype(np.float64))
    elif e.evidence:
        val = val * (1 - 0.5 ** mat(e.evidence, C.c_uint8, e.ld_evidence).astype(np.float64))
    if e.prior:
        val = (1 - e.lambda_) * val + e.lambda_ * mat(e.prior, C.c_double, e.ld_prior)
    jj, rr = np.meshgrid(np.arange(M), np.arange(R), indexing="ij")
    diag = jj == rr + a.diag_offset
    val[diag] = 1.0
    live = np.ones((M, R), dtype=bool) if not sym else (jj <= rr)
    out = mat(a.out_f64, C.c_double, a.ld_out)
    if e.s_old:
        d = np.abs(val - mat(e.s_old, C.c_double, e.ld_s_old))[live]
        d = d[~np.isnan(d)]
        if e.maxdiff and d.size:
            m = _f64(e.maxdiff, 1)
            m[0] = max(m[0], d.max())
    if e.maxoff:
        m = _f64(e.maxoff, 1)
        m[0] = max(m[0], val[live & ~diag].max(initial=0.0))
    if sym:
        up = np.triu(val)
        out[:, :] = up + np.triu(val, 1).T
    else:
        out[:, :] = val
    if a.mirror_out:
        mir = _view(a.mirror_out + 8 * a.mirror_col0, C.c_double, M, R, a.ld_mirror)
        mir[~diag] = val[~diag]


def srk_slice_rows_max_f64(V_ptr, ldv, R, K, zero_diag_offset, ns, planes_ptr, ldp, plane_stride, bound_ptr):
    V = _view(V_ptr, C.c_double, R, K, ldv).copy()
    V[np.isnan(V) | (V < 0)] = 0.0
    if zero_diag_offset >= 0:
        r = np.arange(R)
        ok = r + zero_diag_offset < K
        V[r[ok], r[ok] + zero_diag_offset] = 0.0
    m = V.max(axis=1)
    qmax = 256.0 ** ns - 1
    _f64(bound_ptr, R)[:] = m * ((qmax + 1) / qmax)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.where(m[:, None] > 0, np.rint(V * (qmax / np.where(m > 0, m, 1.0))[:, None]), 0.0).astype(np.int64)
    for s in range(ns):
        out = _view(planes_ptr + s * plane_stride, C.c_uint8, R, ldp, ldp)
        out[:, :K] = ((q >> (8 * (ns - 1 - s))) & 0xFF).astype(np.uint8)
        out[:, K:] = 0


def srk_csr_half_f64(indptr, indices, g, M, row_begin, row_end, X_ptr, ldx, L, OUT_ptr, ldo, epi, K):
    """OUT[c, i] = g[i] * sum_{m in N(i)} X[m, c] for i in [row_begin, row_end), c < L, then the fused
    epilogue when ``epi`` is given.  ``indptr/indices/g`` are host arrays, X has K rows.  Only the
    columns [row_begin, row_end) of OUT are touched (the caller may pass a shifted base)."""
    if row_end <= row_begin or L == 0:
        return
    X = _view(X_ptr, C.c_double, K, L, ldx)
    acc = np.zeros((row_end - row_begin, L))
    for i in range(row_begin, row_end):
        nb = indices[indptr[i]:indptr[i + 1]]
        acc[i - row_begin] = g[i] * X[nb, :].sum(axis=0) if nb.size else 0.0
    val = acc.T                                                    # [L, rows]: (r = c index, column i)
    cols = np.arange(row_begin, row_end)

    def block(ptr, ctype, ld):                                    # rows 0..L, columns row_begin..row_end of a caller matrix
        return _view(ptr + row_begin * C.sizeof(ctype), ctype, L, row_end - row_begin, ld)

    if epi is not None:
        val = epi.coef * val
        if epi.evidence:
            val = val * (1 - 0.5 ** np.minimum(block(epi.evidence, C.c_uint8, epi.ld_evidence).astype(np.float64), 60.0))
        if epi.prior:
            val = (1 - epi.lambda_) * val + epi.lambda_ * block(epi.prior, C.c_double, epi.ld_prior)
        diag = (np.arange(L)[:, None] + epi.diag_offset) == cols[None, :]
        val = np.where(diag, 1.0, val)
        if epi.maxoff:
            m = _f64(epi.maxoff, 1)
            m[0] = max(m[0], val[~diag].max(initial=0.0))
        if epi.s_old and epi.maxdiff:
            d = np.abs(val - block(epi.s_old, C.c_double, epi.ld_s_old))
            d = d[~np.isnan(d)]
            if d.size:
                m = _f64(epi.maxdiff, 1)
                m[0] = max(m[0], d.max())
    block(OUT_ptr, C.c_double, ldo)[:, :] = val


# ------------------------------------------------------------------ fixed-point CSR path (srk_csr_half, U16)
def srk_quantize_rows_u16(V_ptr, ldv, R, K, zero_diag_offset, XT_ptr, ldxt, unit_ptr, qmax, symmetric):
    """unit[r] = max_k V[r, k] / qmax, XT[k, r] = rint(V[r, k] * (1 / unit[r])) clipped (header: srk_quantize_rows_u16)."""
    qmax = 65535.0 if qmax == 0 else float(qmax)
    V = _view(V_ptr, C.c_double, R, K, ldv).copy()
    V[np.isnan(V) | (V < 0)] = 0.0
    if zero_diag_offset >= 0:
        r = np.arange(R)
        ok = r + zero_diag_offset < K
        V[r[ok], r[ok] + zero_diag_offset] = 0.0
    unit = V.max(axis=1) / qmax if K else np.zeros(R)
    _f64(unit_ptr, R)[:] = unit
    with np.errstate(divide="ignore"):
        inv = np.where(unit > 0, 1.0 / np.where(unit > 0, unit, 1.0), 0.0)
    q = np.clip(np.rint(V * inv[:, None]), 0, qmax).astype(np.uint16)
    XT = _view(XT_ptr, C.c_uint16, K, ldxt, ldxt)
    XT[:, :R] = q.T
    XT[:, R:] = 0


def _lists(a, rows):
    """Neighbour list of every row in ``rows``: split bounds when given, else indptr."""
    if a.row_lo:
        lo = np.ctypeslib.as_array((C.c_int64 * a.M).from_address(a.row_lo))
        hi = np.ctypeslib.as_array((C.c_int64 * a.M).from_address(a.row_hi))
    else:
        ptr = np.ctypeslib.as_array((C.c_int64 * (a.M + 1)).from_address(a.indptr))
        lo, hi = ptr[:-1], ptr[1:]
    idx = np.ctypeslib.as_array((C.c_int32 * max(int(hi.max(initial=0)), 1)).from_address(a.indices))
    return [idx[lo[r]:hi[r]] for r in rows]


def srk_csr_half(a: _lib.CsrArgs) -> None:
    """numpy statement of srk_csr_half for what the row-sharded solvers call: U16 FIRST / FINAL / ACCUM / FINISH /
    FINISH_FIRST (transposed store; the symmetric variants belong to the single-GPU solver) and F64 FINAL with
    the evidence taken from ``counts``."""
    u16 = a.elem == _lib.SRK_ELEM_U16
    qmax = 65535.0 if a.qmax == 0 else float(a.qmax)
    L, K = a.L, a.K
    rows = np.arange(a.row_begin, a.row_end)
    if rows.size == 0 or L == 0:
        return
    assert not a.symmetric or a.mode == _lib.SRK_CSR_ACCUM, "symmetric second halves are not emulated"
    acc = None
    if a.accum:
        slots = 1 + max(rows.max(), a.M)                              # enough rows for any slot in use
        acc = _view(a.accum, C.c_uint32, slots, a.ld_accum, a.ld_accum)
    if a.mode in (_lib.SRK_CSR_FIRST, _lib.SRK_CSR_FINAL, _lib.SRK_CSR_ACCUM):
        X = _view(a.X, C.c_uint16 if u16 else C.c_double, K, L, a.ldx)
        X = X.astype(np.int64) if u16 else X
        D = np.stack([X[nb].sum(axis=0) if nb.size else np.zeros(L, dtype=X.dtype) for nb in _lists(a, rows)])
    if a.mode == _lib.SRK_CSR_ACCUM:
        slot = np.ctypeslib.as_array((C.c_int32 * a.M).from_address(a.accum_slot))[rows]
        for t, s in enumerate(slot):
            if s >= 0:
                acc[s, :L] += D[t].astype(np.uint32)
            else:
                acc[-s - 1, :L] = D[t].astype(np.uint32)
        return
    if a.mode in (_lib.SRK_CSR_FINISH, _lib.SRK_CSR_FINISH_FIRST):
        D = acc[rows, :L].astype(np.int64)
    elif u16 and a.accum_slot:                                        # FIRST / FINAL start from the pre-summed pieces
        slot = np.ctypeslib.as_array((C.c_int32 * a.M).from_address(a.accum_slot))[rows]
        D = D + np.where(slot[:, None] >= 0, acc[np.maximum(slot, 0), :L].astype(np.int64), 0)
    assert not u16 or D.max(initial=0) < (1 << 32), "32-bit sums overflow: qmax too large for the degree"

    def block(ptr, ctype, ld):                                        # [L rows r/c, columns row_begin..row_end) of a caller matrix
        return _view(ptr + a.row_begin * C.sizeof(ctype), ctype, L, rows.size, ld)

    unit = _bound(a.in_unit, L) if u16 else None
    if a.mode in (_lib.SRK_CSR_FIRST, _lib.SRK_CSR_FINISH_FIRST):
        assert u16
        ob = _bound(a.out_bound, a.row_end)[rows]
        with np.errstate(divide="ignore"):
            inv = np.where(ob > 0, qmax / np.where(ob > 0, ob, 1.0), 0.0)
        q = np.clip(np.rint((D.astype(np.float64) * unit[None, :]) * inv[:, None]), 0, qmax)
        block(a.OUT, C.c_uint16, a.ldo)[:, :] = q.T.astype(np.uint16)
        return
    # ---- FINAL / FINISH: value at (r, i), r = column of X = output row
    e = a.epi
    g = _f64(a.g, a.row_end)[rows]
    cnt = None
    if a.counts:
        cnt = block(a.counts, C.c_uint32 if a.counts_bits == 32 else C.c_uint16, a.ld_counts).astype(np.int64)
    if u16:
        gc = _f64(a.g_col, L)
        inner = D.T.astype(np.float64) * unit[:, None] + (cnt.astype(np.float64) if a.add_counts else 0.0)
        val = ((g[None, :] * gc[:, None]) * inner) * e.coef
    else:
        val = (D.T * g[None, :]) * e.coef
    if a.use_evidence:
        val = val * (1 - 0.5 ** np.minimum(cnt, 60).astype(np.float64))
    elif e.evidence:
        val = val * (1 - 0.5 ** np.minimum(block(e.evidence, C.c_uint8, e.ld_evidence).astype(np.float64), 60.0))
    if e.prior:
        val = (1 - e.lambda_) * val + e.lambda_ * block(e.prior, C.c_double, e.ld_prior)
    diag = (np.arange(L)[:, None] + e.diag_offset) == rows[None, :]
    val = np.where(diag, 1.0, val)
    if e.maxoff:
        m = _f64(e.maxoff, 1)
        m[0] = max(m[0], val[~diag].max(initial=0.0))
    if e.s_old and e.maxdiff:
        d = np.abs(val - block(e.s_old, C.c_double, e.ld_s_old))
        d = d[~np.isnan(d)]
        if d.size:
            m = _f64(e.maxdiff, 1)
            m[0] = max(m[0], d.max())
    block(a.OUT, C.c_double, a.ldo)[:, :] = val


class Library:
    """Stand-in for the ctypes library object with the two entry points the row-sharded fixed-point solver calls."""

    @staticmethod
    def srk_csr_half(ref, stream):
        srk_csr_half(ref._obj)
        return 0

    @staticmethod
    def srk_quantize_rows_u16(V, ldv, R, K, zd, XT, ldxt, unit, qmax, symmetric, stream):
        srk_quantize_rows_u16(V.value, ldv, R, K, zd, XT.value, ldxt, unit.value, qmax, symmetric)
        return 0
