"""Kernel-level tests of the CSR half-products with the shared-memory gather ring
(srk_csr_half: float64 and uint16 fixed point) and of srk_quantize_rows_u16, through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

from simrank_b200 import _lib, engine, graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return engine.require_cuda()


def _rand_graph(rng, M, K, density, empty_rows=0):
    mask = rng.random((M, K)) < density
    if empty_rows:
        mask[rng.choice(M, size=empty_rows, replace=False)] = False
    return graph.operator_from_edges(*np.nonzero(mask), M, K, rng.random(M) * 0.2 + 0.01), mask.astype(np.int64)


def _u16(t):                      # torch has no uint16 arithmetic: int16 tensors carry the bytes
    return torch.from_numpy(t.astype(np.uint16).view(np.int16))


def _args(dop, elem, mode):
    a = _lib.CsrArgs()
    a.elem, a.mode = elem, mode
    a.indptr, a.indices, a.g = dop.indptr.data_ptr(), dop.indices.data_ptr(), dop.g.data_ptr()
    a.M, a.row_begin, a.row_end = dop.M, 0, dop.M
    return a


@pytest.mark.parametrize("R,K,diag", [(5, 7, -1), (64, 64, 0), (200, 333, 17), (130, 65, -1)])
def test_quantize_rows_u16(dev, R, K, diag):
    rng = np.random.default_rng(R + K)
    V = rng.random((R, K)) * rng.random((R, 1))
    V[R // 2] = 0.0                                        # an all-zero row: unit 0, zeros out
    V[0, :2] = [-1.0, np.nan]
    ldxt = engine._round_up(R, 64)
    Vd = torch.from_numpy(V).to(dev)
    xt = torch.full((K, ldxt), -1, dtype=torch.int16, device=dev)
    unit = torch.zeros(R, dtype=torch.float64, device=dev)
    _lib.check(_lib.load().srk_quantize_rows_u16(engine._ptr(Vd), K, R, K, diag, engine._ptr(xt), ldxt,
                                                 engine._ptr(unit), 65535.0, 0, engine._stream()))
    torch.cuda.synchronize()
    W = np.where(np.isnan(V) | (V < 0), 0.0, V)
    if diag >= 0:
        idx = np.arange(R)
        ok = idx + diag < K
        W[idx[ok], idx[ok] + diag] = 0.0
    u = W.max(axis=1) / 65535.0
    np.testing.assert_array_equal(unit.cpu().numpy(), u)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.clip(np.where(u[:, None] > 0, np.rint(W * (1.0 / u)[:, None]), 0.0), 0, 65535)   # scaled by the reciprocal
    got = xt.cpu().numpy().view(np.uint16)
    np.testing.assert_array_equal(got[:, :R], q.T.astype(np.uint16))
    assert not got[:, R:].any()


@pytest.mark.parametrize("n", [5, 64, 333])
def test_quantize_symmetric_matrix_without_transposition(dev, n):
    rng = np.random.default_rng(n)
    V = rng.random((n, n)) * rng.random((n, 1))
    V = np.triu(V) + np.triu(V, 1).T
    ldxt = engine._round_up(n, 64)
    Vd = torch.from_numpy(V).to(dev)
    outs = []
    for sym in (0, 1):
        xt = torch.full((n, ldxt), -1, dtype=torch.int16, device=dev)
        unit = torch.zeros(n, dtype=torch.float64, device=dev)
        _lib.check(_lib.load().srk_quantize_rows_u16(engine._ptr(Vd), n, n, n, 0, engine._ptr(xt), ldxt,
                                                     engine._ptr(unit), 65535.0, sym, engine._stream()))
        torch.cuda.synchronize()
        outs.append((xt.cpu().numpy(), unit.cpu().numpy()))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("M,K,L,density", [(37, 53, 29, 0.3), (300, 257, 260, 0.05), (129, 130, 515, 0.5),
                                           (64, 2000, 128, 0.02), (1, 1, 1, 1.0)])
@pytest.mark.parametrize("aligned", [True, False])
def test_csr16_first_half(dev, M, K, L, density, aligned):
    """uint16 gather, exact integer sums, re-quantised transposed store -- bulk-copy ring (aligned
    operand) and plain-load fallback (odd leading dimension)."""
    rng = np.random.default_rng(M * 7 + L)
    op, A = _rand_graph(rng, M, K, density, empty_rows=min(2, M - 1))
    dop = engine.DeviceOperator(op, dev)
    ldx = engine._round_up(L, 8) if aligned else engine._round_up(L, 8) + 3
    Xq = rng.integers(0, 65536, (K, L), dtype=np.int64)
    Xq[:, 0] = 65535                                       # a column at the top of the range
    xp = np.zeros((K, ldx), dtype=np.int64)
    xp[:, :L] = Xq
    X = _u16(xp).to(dev)
    unit = rng.random(L) * 1e-6
    ob = rng.random(M) * 0.05 + 1e-4
    ob[M // 3] = 0.0                                       # bound 0: the column is stored as zeros
    ud, od = torch.from_numpy(unit).to(dev), torch.from_numpy(ob).to(dev)
    ldo = engine._round_up(M, 8)
    out = torch.full((L, ldo), -1, dtype=torch.int16, device=dev)
    a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FIRST)
    a.X, a.ldx, a.L, a.OUT, a.ldo = X.data_ptr(), ldx, L, out.data_ptr(), ldo
    a.K = K if M % 2 else 0                                # rows of X given / not given (TMA bounds check)
    a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
    a.out_bound = _lib.RowBound.of(od.data_ptr(), 2.0, 0.0)
    _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    D = A @ Xq                                             # [M, L] exact
    with np.errstate(divide="ignore"):
        inv = np.where(ob > 0, 65535.0 / (2.0 * ob), 0.0)
    want = np.clip(np.rint(D.astype(np.float64) * unit[None, :] * inv[:, None]), 0, 65535).T
    got = out.cpu().numpy().view(np.uint16)[:, :M].astype(np.int64)
    np.testing.assert_array_equal(got, want.astype(np.int64))


def _final_case(rng, dev, n, density, evidence_from_counts):
    op, A = _rand_graph(rng, n, n, density, empty_rows=3)
    dop = engine.DeviceOperator(op, dev)
    Tq = rng.integers(0, 65536, (n, n), dtype=np.int64)
    ldt = engine._round_up(n, 64)
    tp = np.zeros((n, ldt), dtype=np.int64)
    tp[:, :n] = Tq
    cnt = rng.integers(0, 40, (n, n))
    cnt = np.triu(cnt) + np.triu(cnt, 1).T                 # counts / S_old are symmetric in the product
    S_old = rng.random((n, n))
    S_old = np.triu(S_old) + np.triu(S_old, 1).T
    unit = rng.random(n) * 1e-5
    gcol = rng.random(n) * 0.3
    D2 = A @ Tq                                            # [i, r]
    x = op.g[:, None] * gcol[None, :] * (D2 * unit[None, :] + cnt) * 0.8          # [i, r]
    if evidence_from_counts:
        x = x * (1 - 0.5 ** cnt.astype(np.float64))
    return op, dop, tp, ldt, cnt, S_old, unit, gcol, x


@pytest.mark.parametrize("n,density", [(70, 0.3), (300, 0.05), (515, 0.1)])
@pytest.mark.parametrize("evidence", [False, True])
def test_csr16_final_transposed(dev, n, density, evidence):
    rng = np.random.default_rng(n)
    op, dop, tp, ldt, cnt, S_old, unit, gcol, x = _final_case(rng, dev, n, density, evidence)
    r0, L = 16, n - 23                                     # a row block of the output: panel T[:, r0 : r0 + L)
    ld = engine._round_up(n, 16)
    ldx = engine._round_up(L, 8)
    xp = np.zeros((n, ldx), dtype=np.int64)
    xp[:, :L] = tp[:, r0:r0 + L]
    X = _u16(xp).to(dev)
    out = torch.zeros((L, ld), dtype=torch.float64, device=dev)
    out[:, :n] = torch.from_numpy(S_old[r0:r0 + L])
    c16 = torch.from_numpy(np.ascontiguousarray(cnt[r0:r0 + L]).astype(np.uint16).view(np.int16)).to(dev)
    ud, gd = torch.from_numpy(unit[r0:r0 + L].copy()).to(dev), torch.from_numpy(gcol[r0:r0 + L].copy()).to(dev)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FINAL)
    a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = X.data_ptr(), ldx, L, n, out.data_ptr(), ld
    a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
    a.g_col = gd.data_ptr()
    a.counts, a.ld_counts, a.counts_bits, a.add_counts, a.use_evidence = c16.data_ptr(), n, 16, 1, int(evidence)
    a.epi.coef = 0.8
    a.epi.s_old, a.epi.ld_s_old = out.data_ptr(), ld
    a.epi.maxdiff, a.epi.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    a.epi.diag_offset = r0
    _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    want = x.T[r0:r0 + L].copy()                           # [r, i]
    want[np.arange(L), r0 + np.arange(L)] = 1.0
    got = out[:, :n].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-300)
    md, mo = scal.tolist()
    assert md == np.abs(got - S_old[r0:r0 + L]).max()
    off = got.copy()
    off[np.arange(L), r0 + np.arange(L)] = 0.0
    assert mo == off.max()


@pytest.mark.parametrize("n,density", [(70, 0.3), (300, 0.05), (515, 0.1), (1100, 0.02)])
@pytest.mark.parametrize("evidence", [False, True])
def test_csr16_final_symmetric(dev, n, density, evidence):
    """Pairs r >= i are computed once (from row i of the graph) and mirrored: the result is the
    upper triangle of the full product, reflected."""
    rng = np.random.default_rng(n + 1)
    op, dop, tp, ldt, cnt, S_old, unit, gcol, x = _final_case(rng, dev, n, density, evidence)
    ld = engine._round_up(n, 16)
    X = _u16(tp).to(dev)
    out = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    out[:, :n] = torch.from_numpy(S_old)
    c16 = torch.from_numpy(cnt.astype(np.uint16).view(np.int16)).to(dev)
    ud, gd = torch.from_numpy(unit).to(dev), torch.from_numpy(gcol).to(dev)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FINAL)
    a.symmetric = 1
    a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = X.data_ptr(), ldt, n, n, out.data_ptr(), ld
    a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
    a.g_col = gd.data_ptr()
    a.counts, a.ld_counts, a.counts_bits, a.add_counts, a.use_evidence = c16.data_ptr(), n, 16, 1, int(evidence)
    a.epi.coef = 0.8
    a.epi.s_old, a.epi.ld_s_old = out.data_ptr(), ld
    a.epi.maxdiff, a.epi.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    want = np.triu(x, 1)                                   # x[i, r] for r > i
    want = want + want.T
    np.fill_diagonal(want, 1.0)
    got = out[:, :n].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-300)
    assert np.array_equal(got, got.T)
    md, mo = scal.tolist()
    assert md == np.abs(got - S_old).max()
    off = got.copy()
    np.fill_diagonal(off, 0.0)
    assert mo == off.max()


@pytest.mark.parametrize("n,density", [(203, 0.1), (700, 0.03)])
def test_csr_f64_final_symmetric_matches_the_full_product(dev, n, density):
    rng = np.random.default_rng(n + 2)
    op, A = _rand_graph(rng, n, n, density, empty_rows=4)
    dop = engine.DeviceOperator(op, dev)
    G = op.to_dense()
    S0 = rng.random((n, n))
    S0 = (S0 + S0.T) / 2
    T = (G @ S0).T                                         # exact first half of a symmetric S
    ld = engine._round_up(n, 16)
    Td = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    Td[:, :n] = torch.from_numpy(T)
    out = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    out[:, :n] = torch.from_numpy(S0)
    ev = rng.integers(0, 60, (n, n))
    ev = (np.triu(ev) + np.triu(ev, 1).T).astype(np.uint8)
    evd = torch.zeros((n, ld), dtype=torch.uint8, device=dev)
    evd[:, :n] = torch.from_numpy(ev)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    a = _args(dop, _lib.SRK_ELEM_F64, _lib.SRK_CSR_FINAL)
    a.symmetric = 1
    a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = Td.data_ptr(), ld, n, n, out.data_ptr(), ld
    a.epi.coef = 0.6
    a.epi.evidence, a.epi.ld_evidence = evd.data_ptr(), ld
    a.epi.s_old, a.epi.ld_s_old = out.data_ptr(), ld
    a.epi.maxdiff, a.epi.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    want = (1 - 0.5 ** ev.astype(np.float64)) * 0.6 * (G @ S0 @ G.T)
    np.fill_diagonal(want, 1.0)
    got = out[:, :n].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    assert np.array_equal(got, got.T)
    md, mo = scal.tolist()
    assert md == np.abs(got - S0).max()


def test_csr_half_rejects_bad_arguments(dev):
    rng = np.random.default_rng(0)
    op, _ = _rand_graph(rng, 20, 20, 0.3)
    dop = engine.DeviceOperator(op, dev)
    lib = _lib.load()
    X = torch.zeros((20, 24), dtype=torch.float64, device=dev)
    a = _args(dop, 7, _lib.SRK_CSR_FIRST)
    a.X, a.ldx, a.L, a.OUT, a.ldo = X.data_ptr(), 24, 20, X.data_ptr(), 24
    assert lib.srk_csr_half(C.byref(a), engine._stream()) == -1 and b"elem" in lib.srk_last_error()
    a = _args(dop, _lib.SRK_ELEM_F64, _lib.SRK_CSR_FINAL)
    a.symmetric, a.row_end = 1, 10
    a.X, a.ldx, a.L, a.OUT, a.ldo = X.data_ptr(), 24, 20, X.data_ptr(), 24
    assert lib.srk_csr_half(C.byref(a), engine._stream()) == -1 and b"symmetric" in lib.srk_last_error()
    a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FINAL)
    a.X, a.ldx, a.L, a.OUT, a.ldo = X.data_ptr(), 24, 20, X.data_ptr(), 24
    assert lib.srk_csr_half(C.byref(a), engine._stream()) == -1 and b"g_col" in lib.srk_last_error()


def test_csr16_row_degree_above_65536_uses_a_coarser_range(dev):
    """A row with more than 65536 neighbours (a popular item of a ratings graph): the 32-bit sums stay
    exact because every matrix it gathers from is held with qmax = floor((2^32 - 1) / degree) levels."""
    rng = np.random.default_rng(9)
    M, K, L, big = 40, 70001, 96, 70000
    rows = np.concatenate([np.full(big, 7), rng.integers(0, M, 500)])
    cols = np.concatenate([np.arange(big), rng.integers(0, K, 500)])
    keep = np.unique(rows * K + cols, return_index=True)[1]
    op = graph.operator_from_edges(rows[keep], cols[keep], M, K)
    qmax = engine.gather_qmax(op.deg)
    assert op.deg.max() >= big and qmax == float(((1 << 32) - 1) // int(op.deg.max())) < 65535
    dop = engine.DeviceOperator(op, dev)
    Xq = rng.integers(0, int(qmax) + 1, (K, L), dtype=np.int64)
    Xq[:, 3] = int(qmax)                                   # the worst case: every neighbour at the top of the range
    X = _u16(Xq).to(dev)
    unit = rng.random(L) * 1e-7
    ob = (op.deg * 1e-7 * qmax + 1e-12).astype(np.float64)  # bound >= deg * max value: nothing clips
    ud, od = torch.from_numpy(unit).to(dev), torch.from_numpy(ob).to(dev)
    ldo = engine._round_up(M, 8)
    out = torch.full((L, ldo), -1, dtype=torch.int16, device=dev)
    a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FIRST)
    a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = X.data_ptr(), L, L, K, out.data_ptr(), ldo
    a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
    a.out_bound = _lib.RowBound.of(od.data_ptr(), 1.0, 0.0)
    a.qmax = qmax
    _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    A = np.zeros((M, K), dtype=np.int64)
    A[rows[keep], cols[keep]] = 1
    D = A @ Xq
    assert D.max() > (1 << 31)                             # the sums really use the whole 32 bits
    want = np.clip(np.rint(D.astype(np.float64) * unit[None, :] * (qmax / ob)[:, None]), 0, qmax).T
    got = out.cpu().numpy().view(np.uint16)[:, :M].astype(np.int64)
    np.testing.assert_array_equal(got, want.astype(np.int64))


# ------------------------------------------------------------------ split neighbour lists (SRK_CSR_ACCUM)
def _skewed_graph(rng, M, K, hubs):
    """A few rows with hundreds of neighbours among rows with a handful (a ratings graph in small)."""
    mask = rng.random((M, K)) < 0.02
    for r, d in hubs:
        mask[r] = False
        mask[r, rng.choice(K, size=d, replace=False)] = True
    return graph.operator_from_edges(*np.nonzero(mask), M, K, rng.random(M) * 0.2 + 0.01), mask.astype(np.int64)


def _run_split(dop, a, K, min_deg, piece, ranges):
    sp = engine.ListSplit(dop.indptr, dop.indices, K, min_deg, piece, ranges)
    sp.accumulate(_lib.load(), a.indices, a.X, a.ldx, a.L, a.K, a.qmax)
    sp.attach(a)
    _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    return sp


@pytest.mark.parametrize("piece,ranges", [(8, 1), (64, 3), (1024, 2)])
def test_split_lists_first_half_is_bit_identical(dev, piece, ranges):
    rng = np.random.default_rng(piece)
    M, K, L = 150, 900, 700                                # two column panels, the second one partial
    op, A = _skewed_graph(rng, M, K, [(0, 900), (17, 333), (149, 512), (60, 100)])
    dop = engine.DeviceOperator(op, dev)
    ldx = engine._round_up(L, 8)
    Xq = rng.integers(0, 65536, (K, ldx), dtype=np.int64)
    X = _u16(Xq).to(dev)
    ud = torch.from_numpy(rng.random(L) * 1e-6).to(dev)
    od = torch.from_numpy(op.deg * 1e-6 * 65535 + 1e-9).to(dev)
    ldo = engine._round_up(M, 8)
    outs = []
    for split in (False, True):
        out = torch.full((L, ldo), -1, dtype=torch.int16, device=dev)
        a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FIRST)
        a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = X.data_ptr(), ldx, L, K, out.data_ptr(), ldo
        a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
        a.out_bound = _lib.RowBound.of(od.data_ptr(), 1.0, 0.0)
        if split:
            sp = _run_split(dop, a, K, 100, piece, ranges)
            assert sp.rows == 4 and sp.pieces >= 4 * ranges
        else:
            _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
            torch.cuda.synchronize()
        outs.append(out.cpu().numpy().view(np.uint16)[:, :M])
    np.testing.assert_array_equal(outs[0], outs[1])
    D = A @ Xq[:, :L]
    want = np.clip(np.rint(D * ud.cpu().numpy()[None, :] * (65535.0 / od.cpu().numpy())[:, None]), 0, 65535).T
    np.testing.assert_array_equal(outs[1].astype(np.int64), want.astype(np.int64))


@pytest.mark.parametrize("symmetric", [0, 1])
def test_split_lists_second_half_is_bit_identical(dev, symmetric):
    rng = np.random.default_rng(40 + symmetric)
    n = 600
    op, A = _skewed_graph(rng, n, n, [(3, 600), (200, 333), (599, 257)])
    dop = engine.DeviceOperator(op, dev)
    ldt, ld = engine._round_up(n, 64), engine._round_up(n, 16)
    X = _u16(rng.integers(0, 65536, (n, ldt), dtype=np.int64)).to(dev)
    cnt = rng.integers(0, 40, (n, n))
    c16 = torch.from_numpy((np.triu(cnt) + np.triu(cnt, 1).T).astype(np.uint16).view(np.int16)).to(dev)
    S_old = rng.random((n, n))
    S_old = np.triu(S_old) + np.triu(S_old, 1).T
    ud, gd = torch.from_numpy(rng.random(n) * 1e-5).to(dev), torch.from_numpy(rng.random(n) * 0.3).to(dev)
    res = []
    for split in (False, True):
        out = torch.zeros((n, ld), dtype=torch.float64, device=dev)
        out[:, :n] = torch.from_numpy(S_old)
        scal = torch.zeros(2, dtype=torch.float64, device=dev)
        a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FINAL)
        a.symmetric = symmetric
        a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = X.data_ptr(), ldt, n, n, out.data_ptr(), ld
        a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
        a.g_col = gd.data_ptr()
        a.counts, a.ld_counts, a.counts_bits, a.add_counts, a.use_evidence = c16.data_ptr(), n, 16, 1, 1
        a.epi.coef = 0.8
        a.epi.s_old, a.epi.ld_s_old = out.data_ptr(), ld
        a.epi.maxdiff, a.epi.maxoff = scal.data_ptr(), scal.data_ptr() + 8
        if split:
            _run_split(dop, a, n, 200, 32, 2)
        else:
            _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
            torch.cuda.synchronize()
        res.append((out[:, :n].cpu().numpy(), scal.tolist()))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1]


def test_split_lists_through_the_solver(dev, monkeypatch):
    """The csr16 solver with hub rows pre-summed gives the matrix of the unsplit solver, bit for bit."""
    rng = np.random.default_rng(77)
    n = 1200
    op, _ = _skewed_graph(rng, n, n, [(5, 1100), (6, 900), (700, 640)])
    mats = []
    monkeypatch.setenv("SRK_CSR_VIA_ACCUM", "0")            # the fused launches, with and without the hub split
    for min_deg in ("0", "256"):
        monkeypatch.setenv("SRK_SPLIT_MIN", min_deg)
        monkeypatch.setenv("SRK_SPLIT_PIECE", "100")
        monkeypatch.setenv("SRK_SPLIT_RANGE_MB", "0.5")
        solver = engine.DirectedSolver(engine.DeviceOperator(op, dev), 0.8, mode="csr16")
        assert (solver.half.split is not None) == (min_deg != "0")
        for _ in range(4):
            solver.step()
        mats.append(solver.S.cpu().numpy().copy())
    np.testing.assert_array_equal(mats[0], mats[1])


def test_accum_mode_rejects_bad_arguments(dev):
    rng = np.random.default_rng(1)
    op, _ = _rand_graph(rng, 20, 20, 0.3)
    dop = engine.DeviceOperator(op, dev)
    lib = _lib.load()
    X = torch.zeros((20, 24), dtype=torch.int16, device=dev)
    a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_ACCUM)
    a.X, a.ldx, a.L = X.data_ptr(), 24, 20
    assert lib.srk_csr_half(C.byref(a), engine._stream()) == -1 and b"SRK_CSR_ACCUM needs" in lib.srk_last_error()
    acc = torch.zeros((4, 512), dtype=torch.int32, device=dev)
    a = _args(dop, _lib.SRK_ELEM_F64, _lib.SRK_CSR_FIRST)
    a.X, a.ldx, a.L, a.OUT, a.ldo = X.data_ptr(), 24, 20, X.data_ptr(), 24
    a.accum, a.ld_accum, a.accum_slot = acc.data_ptr(), 512, acc.data_ptr()
    assert lib.srk_csr_half(C.byref(a), engine._stream()) == -1 and b"fixed-point" in lib.srk_last_error()
    a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FIRST)
    a.X, a.ldx, a.L, a.OUT, a.ldo = X.data_ptr(), 24, 20, X.data_ptr(), 24
    a.accum, a.ld_accum, a.accum_slot = acc.data_ptr(), 500, acc.data_ptr()
    assert lib.srk_csr_half(C.byref(a), engine._stream()) == -1 and b"ld_accum" in lib.srk_last_error()


@pytest.mark.parametrize("piece,ranges", [(16, 1), (64, 3)])
def test_second_half_as_accum_plus_finish_is_bit_identical(dev, piece, ranges):
    """SRK_CSR_ACCUM over EVERY neighbour list (single-piece lists stored, the others added) followed by
    SRK_CSR_FINISH gives the matrix of one SRK_CSR_FINAL launch, bit for bit (transposed store, a row block
    of the output with its diag_offset)."""
    rng = np.random.default_rng(piece)
    n, r0, L = 700, 32, 600
    op, A = _skewed_graph(rng, n, n, [(3, 650), (200, 333), (699, 257), (11, 0), (12, 1)])
    dop = engine.DeviceOperator(op, dev)
    ldx, ld = engine._round_up(L, 8), engine._round_up(n, 16)
    X = _u16(rng.integers(0, 65536, (n, ldx), dtype=np.int64)).to(dev)
    c16 = torch.from_numpy(rng.integers(0, 40, (L, n)).astype(np.uint16).view(np.int16)).to(dev)
    S_old = rng.random((L, n))
    ud, gd = torch.from_numpy(rng.random(L) * 1e-5).to(dev), torch.from_numpy(rng.random(L) * 0.3).to(dev)
    res = []
    for via_accum in (False, True):
        out = torch.zeros((L, ld), dtype=torch.float64, device=dev)
        out[:, :n] = torch.from_numpy(S_old)
        scal = torch.zeros(2, dtype=torch.float64, device=dev)
        a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FINAL)
        a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = X.data_ptr(), ldx, L, n, out.data_ptr(), ld
        a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
        a.g_col = gd.data_ptr()
        a.counts, a.ld_counts, a.counts_bits, a.add_counts, a.use_evidence = c16.data_ptr(), n, 16, 1, 1
        a.epi.coef = 0.8
        a.epi.s_old, a.epi.ld_s_old = out.data_ptr(), ld
        a.epi.maxdiff, a.epi.maxoff = scal.data_ptr(), scal.data_ptr() + 8
        a.epi.diag_offset = r0
        if via_accum:
            sp = engine.ListSplit(dop.indptr, dop.indices, n, 1, piece, ranges, all_rows=True)
            assert sp.rows == n and (sp.piece_slot < 0).any() and (sp.piece_slot >= 0).any()
            sp.accumulate(_lib.load(), a.indices, a.X, a.ldx, a.L, a.K, 65535.0)
            a.mode, a.accum, a.ld_accum = _lib.SRK_CSR_FINISH, sp._accum.data_ptr(), sp._accum.shape[1]
        _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
        torch.cuda.synchronize()
        res.append((out[:, :n].cpu().numpy(), scal.tolist()))
    np.testing.assert_array_equal(res[0][0], res[1][0])
    assert res[0][1] == res[1][1]
    assert res[0][0][5, r0 + 5] == 1.0


@pytest.mark.parametrize("piece,ranges", [(16, 1), (100, 3)])
def test_first_half_as_accum_plus_finish_is_bit_identical(dev, piece, ranges):
    rng = np.random.default_rng(piece + 5)
    M, K, L = 150, 900, 700
    op, A = _skewed_graph(rng, M, K, [(0, 900), (17, 333), (149, 512), (60, 0)])
    dop = engine.DeviceOperator(op, dev)
    ldx = engine._round_up(L, 8)
    X = _u16(rng.integers(0, 65536, (K, ldx), dtype=np.int64)).to(dev)
    ud = torch.from_numpy(rng.random(L) * 1e-6).to(dev)
    ob = op.deg * 1e-6 * 65535 + 1e-9
    ob[7] = 0.0                                            # bound 0: the column is stored as zeros
    od = torch.from_numpy(ob).to(dev)
    ldo = engine._round_up(M, 8)
    outs = []
    for via_accum in (False, True):
        out = torch.full((L, ldo), -1, dtype=torch.int16, device=dev)
        a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FIRST)
        a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = X.data_ptr(), ldx, L, K, out.data_ptr(), ldo
        a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
        a.out_bound = _lib.RowBound.of(od.data_ptr(), 1.0, 0.0)
        if via_accum:
            sp = engine.ListSplit(dop.indptr, dop.indices, K, 1, piece, ranges, all_rows=True)
            sp.accumulate(_lib.load(), a.indices, a.X, a.ldx, a.L, a.K, 65535.0)
            a.mode, a.accum, a.ld_accum = _lib.SRK_CSR_FINISH_FIRST, sp._accum.data_ptr(), sp._accum.shape[1]
        _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy().view(np.uint16)[:, :M].copy())
    np.testing.assert_array_equal(outs[0], outs[1])
    assert outs[1].any()


@pytest.mark.parametrize("n,density", [(70, 0.3), (300, 0.05), (515, 0.1), (1100, 0.02)])
@pytest.mark.parametrize("evidence", [False, True])
def test_symmetric_second_half_as_accum_plus_finish(dev, n, density, evidence):
    """ACCUM with upper_only (a row's pieces skip the panels left of its diagonal) + the symmetric FINISH:
    the upper triangle of the full product, reflected -- the reference of test_csr16_final_symmetric."""
    rng = np.random.default_rng(n + 1)
    op, dop, tp, ldt, cnt, S_old, unit, gcol, x = _final_case(rng, dev, n, density, evidence)
    ld = engine._round_up(n, 16)
    X = _u16(tp).to(dev)
    out = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    out[:, :n] = torch.from_numpy(S_old)
    c16 = torch.from_numpy(cnt.astype(np.uint16).view(np.int16)).to(dev)
    ud, gd = torch.from_numpy(unit).to(dev), torch.from_numpy(gcol).to(dev)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    a = _args(dop, _lib.SRK_ELEM_U16, _lib.SRK_CSR_FINISH)
    a.symmetric = 1
    a.X, a.ldx, a.L, a.K, a.OUT, a.ldo = X.data_ptr(), ldt, n, n, out.data_ptr(), ld
    a.in_unit = _lib.RowBound.of(ud.data_ptr(), 1.0, 0.0)
    a.g_col = gd.data_ptr()
    a.counts, a.ld_counts, a.counts_bits, a.add_counts, a.use_evidence = c16.data_ptr(), n, 16, 1, int(evidence)
    a.epi.coef = 0.8
    a.epi.s_old, a.epi.ld_s_old = out.data_ptr(), ld
    a.epi.maxdiff, a.epi.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    sp = engine.ListSplit(dop.indptr, dop.indices, n, 1, 24, 2, all_rows=True)
    sp._accum = torch.full((n, engine._round_up(n, 512)), -1, dtype=torch.int32, device=dev)   # skipped panels stay garbage
    sp.accumulate(_lib.load(), a.indices, a.X, a.ldx, a.L, a.K, 65535.0, upper=True)
    a.accum, a.ld_accum = sp._accum.data_ptr(), sp._accum.shape[1]
    _lib.check(_lib.load().srk_csr_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    want = np.triu(x, 1)
    want = want + want.T
    np.fill_diagonal(want, 1.0)
    got = out[:, :n].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-300)
    assert np.array_equal(got, got.T)
    md, mo = scal.tolist()
    assert md == np.abs(got - S_old).max()
    off = got.copy()
    np.fill_diagonal(off, 0.0)
    assert mo == off.max()


def test_csr16_solver_via_accum_matches_the_fused_launches(dev, monkeypatch):
    rng = np.random.default_rng(78)
    n = 1300
    op, _ = _skewed_graph(rng, n, n, [(5, 1100), (6, 900), (700, 640), (9, 0)])
    mats = []
    for via in ("0", "1"):
        monkeypatch.setenv("SRK_CSR_VIA_ACCUM", via)
        solver = engine.DirectedSolver(engine.DeviceOperator(op, dev), 0.8, mode="csr16")
        assert (solver.half.split_all is not None) == (via == "1")
        for _ in range(5):
            d = solver.step()
        mats.append((solver.S.cpu().numpy().copy(), d))
    a, b = mats[0][0], mats[1][0]
    assert np.array_equal(b, b.T) and (np.diag(b) == 1.0).all()
    # the same integer sums; the float64 element is rounded in another order, which can move a value across a
    # quantisation step of the next update (one step of 2^-16 of a row maximum, contracted by C)
    assert np.abs(a - b).max() <= 5e-8
    assert abs(mats[0][1] - mats[1][1]) <= 5e-8
