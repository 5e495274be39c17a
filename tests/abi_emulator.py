"""Host-memory emulator of ``srk_i8_half`` (include/simrank_b200.h) in numpy.  TEST ONLY.

It interprets the very ``srk_i8_args`` struct the product passes to the CUDA library, but on
CPU tensors, so the multi-rank host logic (shard offsets, send/receive block layout, K-blocked
operands, bounds) can run under gloo without a GPU.  It is never imported by the package."""
import ctypes as C

import numpy as np

from simrank_b200 import _lib


def _bytes(ptr, n):
    return np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr))


def _f64(ptr, n):
    return np.ctypeslib.as_array((C.c_double * n).from_address(ptr))


def _bound(rb, n):
    if rb.vec:
        return _f64(rb.vec, n) * rb.mul + rb.add
    return np.full(n, rb.add)


def _read_planes(a):
    ns, R, K = a.ns, a.R, a.K
    q = np.zeros((R, K), dtype=np.int64)
    for s in range(ns):
        if a.in_kblock > 0:
            kb = a.in_kblock
            plane = np.zeros((R, K), dtype=np.int64)
            for b in range(K // kb):
                base = a.in_planes + b * a.in_kblock_stride + s * a.in_plane_stride
                raw = _bytes(base, (R - 1) * a.ld_in + kb)
                blk = np.lib.stride_tricks.as_strided(raw, (R, kb), (a.ld_in, 1))
                plane[:, b * kb:(b + 1) * kb] = blk
        else:
            raw = _bytes(a.in_planes + s * a.in_plane_stride, (R - 1) * a.ld_in + K)
            plane = np.lib.stride_tricks.as_strided(raw, (R, K), (a.ld_in, 1)).astype(np.int64)
        q += plane << (8 * (ns - 1 - s))
    return q


def _a8(a):
    raw = _bytes(a.A8, (a.N - 1) * a.lda + a.K)
    return np.lib.stride_tricks.as_strided(raw, (a.N, a.K), (a.lda, 1)).astype(np.int64)


def _write_planes(base, plane_stride, ld, ns, rows, cols, q):
    for s in range(ns):
        raw = _bytes(base + s * plane_stride, (rows - 1) * ld + cols)
        out = np.lib.stride_tricks.as_strided(raw, (rows, cols), (ld, 1))
        out[:, :] = ((q >> (8 * (ns - 1 - s))) & 0xFF).astype(np.uint8)


def srk_i8_half(a: _lib.I8Args) -> None:
    ns, R, N, K = a.ns, a.R, a.N, a.K
    kq = float(256 ** ns)
    A = _a8(a)
    D = _read_planes(a) @ A.T                                    # exact integers
    if a.mode == _lib.SRK_I8_COUNTS:
        _write_planes(a.out_planes, 0, a.ld_outp, 1, R, N, np.minimum(D, 255))
        return
    inb = _bound(a.in_rowbound, R)
    if a.mode == _lib.SRK_I8_MID:
        U = D * (inb / kq)[:, None]
        if a.unit_diag:
            U = U + A[:, a.diag_offset:a.diag_offset + R].T
        outb = _bound(a.out_rowbound, N)
        scale = np.where(outb > 0, kq / np.where(outb > 0, outb, 1.0), 0.0)
        q = np.clip(np.rint(U * scale[None, :]), 0, kq - 1).astype(np.int64)
        _write_planes(a.out_planes, a.out_plane_stride, a.ld_outp, ns, N, R, q.T.copy())
        return
    e = a.epi
    g_row, g_col = _f64(a.g_row, R), _f64(a.g_col, N)
    val = D * (inb / kq * g_row * e.coef)[:, None] * g_col[None, :]
    if e.evidence:
        raw = _bytes(e.evidence, (R - 1) * e.ld_evidence + N)
        cnt = np.lib.stride_tricks.as_strided(raw, (R, N), (e.ld_evidence, 1)).astype(np.int64)
        val = val * (1 - 0.5 ** cnt)
    if e.prior:
        raw = _f64(e.prior, (R - 1) * e.ld_prior + N)
        pr = np.lib.stride_tricks.as_strided(raw, (R, N), (e.ld_prior * 8, 8))
        val = (1 - e.lambda_) * val + e.lambda_ * pr
    rr = np.arange(R)
    dj = rr + a.diag_offset
    on = dj < N
    val[rr[on], dj[on]] = 1.0
    off = val.copy()
    off[rr[on], dj[on]] = 0.0
    raw = _f64(a.out_f64, (R - 1) * a.ld_out + N)
    out = np.lib.stride_tricks.as_strided(raw, (R, N), (a.ld_out * 8, 8))
    if e.s_old:
        raw = _f64(e.s_old, (R - 1) * e.ld_s_old + N)
        old = np.lib.stride_tricks.as_strided(raw, (R, N), (e.ld_s_old * 8, 8))
        d = np.abs(val - old)
        d = d[~np.isnan(d)]
        if e.maxdiff and d.size:
            m = _f64(e.maxdiff, 1)
            m[0] = max(m[0], d.max())
    if e.maxoff:
        m = _f64(e.maxoff, 1)
        m[0] = max(m[0], off.max(initial=0.0))
    out[:, :] = val
    if a.out_planes:
        outb = _bound(a.out_rowbound, R)
        scale = np.where(outb > 0, kq / np.where(outb > 0, outb, 1.0), 0.0)
        q = np.clip(np.rint(off * scale[:, None]), 0, kq - 1).astype(np.int64)
        _write_planes(a.out_planes, a.out_plane_stride, a.ld_outp, ns, R, N, q)
