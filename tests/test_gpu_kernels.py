"""Kernel-level parity through the C ABI (run on the B200 box: ``pytest -m gpu``).

Integer / index outputs are compared bit-exactly, float64 outputs to 1e-12 relative (the only
differences allowed are summation order and FMA contraction)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import simrank_oracle as orc
from simrank_b200 import _lib, engine, graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return engine.require_cuda()


def _rand_op(rng, M, K, density, weighted=False, empty_rows=0):
    mask = rng.random((M, K)) < density
    if empty_rows:
        mask[rng.choice(M, empty_rows, replace=False)] = False
    rows, cols = np.nonzero(mask)
    g = None
    if weighted:
        g = rng.random(M) * 0.2 + 0.01
    return graph.operator_from_edges(rows, cols, M, K, g), mask


def _csr_half(dop, X, L, final=None, ldo=None):
    lib = _lib.load()
    ldo = ldo or engine._round_up(dop.M, 16)
    out = torch.full((L, ldo), -7.0, dtype=torch.float64, device=X.device)
    rc = lib.srk_csr_half_f64(engine._ptr(dop.indptr), engine._ptr(dop.indices), engine._ptr(dop.g), dop.M, 0,
                              dop.M, engine._ptr(X), X.stride(0), L, engine._ptr(out), ldo,
                              C.byref(final) if final is not None else None, engine._stream())
    _lib.check(rc)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,K,L,density", [(37, 53, 29, 0.3), (300, 257, 260, 0.05), (1, 1, 1, 1.0),
                                           (129, 130, 131, 0.5), (64, 2000, 128, 0.02)])
def test_csr_half_plain(dev, M, K, L, density):
    rng = np.random.default_rng(M * 1000 + K)
    op, _ = _rand_op(rng, M, K, density, weighted=True, empty_rows=min(3, M - 1))
    dop = engine.DeviceOperator(op, dev)
    Xh = rng.random((K, L))
    X = torch.from_numpy(Xh).to(dev)
    out = _csr_half(dop, X, L)[:, :M].cpu().numpy()
    np.testing.assert_allclose(out, (op.to_dense() @ Xh).T, rtol=1e-13, atol=1e-300)


def test_csr_half_final_epilogue(dev):
    rng = np.random.default_rng(5)
    n = 203
    op, mask = _rand_op(rng, n, n, 0.1, weighted=True, empty_rows=4)
    dop = engine.DeviceOperator(op, dev)
    G = op.to_dense()
    Th = rng.random((n, n))
    S_old = rng.random((n, n))
    prior = rng.random((n, n))
    cnt = rng.integers(0, 70, (n, n)).astype(np.uint8)
    ld = engine._round_up(n, 16)
    S_dev = torch.zeros((n, ld), dtype=torch.float64, device=dev)
    S_dev[:, :n] = torch.from_numpy(S_old)
    ev = torch.zeros((n, ld), dtype=torch.uint8, device=dev)
    ev[:, :n] = torch.from_numpy(cnt)
    pr = torch.from_numpy(prior).to(dev)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    e = _lib.Epilogue()
    e.coef, e.lambda_ = 0.8, 0.3
    e.evidence, e.ld_evidence = ev.data_ptr(), ld
    e.prior, e.ld_prior = pr.data_ptr(), n
    e.s_old, e.ld_s_old = S_dev.data_ptr(), ld
    e.maxdiff, e.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    lib = _lib.load()
    T = torch.from_numpy(Th).to(dev)
    _lib.check(lib.srk_csr_half_f64(engine._ptr(dop.indptr), engine._ptr(dop.indices), engine._ptr(dop.g), n, 0, n,
                                    engine._ptr(T), n, n, engine._ptr(S_dev), ld, C.byref(e), engine._stream()))
    torch.cuda.synchronize()
    want = (1 - 0.3) * (1 - 0.5 ** cnt.astype(np.int64)) * 0.8 * (G @ Th).T + 0.3 * prior
    np.fill_diagonal(want, 1.0)
    got = S_dev[:, :n].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    md, mo = scal.tolist()
    assert md == np.abs(got - S_old).max()
    off = got.copy()
    np.fill_diagonal(off, 0)
    assert mo == off.max()


@pytest.mark.parametrize("M,K,density", [(60, 90, 0.3), (500, 300, 0.4), (257, 1000, 0.7)])
def test_evidence_counts_csr(dev, M, K, density):
    rng = np.random.default_rng(M)
    op, mask = _rand_op(rng, M, K, density, weighted=True)
    op.g[: M // 7] = 0.0                                  # rows where G > 0 is False
    dop = engine.DeviceOperator(op, dev)
    cnt = dop.evidence_counts("csr")[:, :M].cpu().numpy()
    A = (op.to_dense() > 0).astype(np.int64)
    np.testing.assert_array_equal(cnt, np.minimum(A @ A.T, 255).astype(np.uint8))


def test_row_spread_general_values(dev):
    rng = np.random.default_rng(2)
    op, mask = _rand_op(rng, 80, 60, 0.3, empty_rows=3)
    dop = engine.DeviceOperator(op, dev)
    vals_h = rng.random(op.nnz)
    vals_h[rng.random(op.nnz) < 0.1] = 0.0               # explicit zeros are skipped (replace(0, nan))
    G = np.zeros((80, 60))
    G[np.repeat(np.arange(80), op.deg), op.indices] = vals_h
    got = dop.row_spread(torch.from_numpy(vals_h).to(dev)).cpu().numpy()
    np.testing.assert_allclose(got, orc.spread(G), rtol=1e-14)
    # reference-built graphs: rows constant on their support -> spread is exactly 1
    assert np.all(dop.row_spread().cpu().numpy() == 1.0)


def test_csr_to_dense_and_slices(dev):
    rng = np.random.default_rng(3)
    op, mask = _rand_op(rng, 70, 200, 0.2)
    dop = engine.DeviceOperator(op, dev)
    a8 = dop.dense_u8().cpu().numpy()
    np.testing.assert_array_equal(a8[:, :200], mask.astype(np.uint8))
    assert not a8[:, 200:].any()
    for ns in (1, 2, 3, 4):
        R, K = 45, 77
        V = rng.random((R, K))
        bvec = rng.random(R) + 0.5
        V *= (bvec * 0.7 + 0.1)[:, None] * 0.999
        ldp = 128
        planes = torch.zeros((ns, R, ldp), dtype=torch.uint8, device=dev)
        Vd = torch.from_numpy(V).to(dev)
        bd = torch.from_numpy(bvec).to(dev)
        rb = _lib.RowBound.of(bd.data_ptr(), 0.7, 0.1)
        _lib.check(_lib.load().srk_slice_rows_f64(engine._ptr(Vd), K, R, K, C.byref(rb), 0, ns, engine._ptr(planes),
                                                  ldp, R * ldp, engine._stream()))
        p = planes.cpu().numpy().astype(np.int64)
        q = sum(p[s] << (8 * (ns - 1 - s)) for s in range(ns))[:, :K]
        bound = bvec * 0.7 + 0.1
        want = np.minimum(np.rint(V * ((256.0 ** ns) / bound)[:, None]), 256.0 ** ns - 1).astype(np.int64)
        want[np.arange(R), np.arange(R)] = 0              # zero_diag_offset = 0
        assert np.abs(q - want).max() <= 1               # FMA contraction may move a tie by one step
        back = q * (bound / 256.0 ** ns)[:, None]
        Vz = V.copy()
        Vz[np.arange(R), np.arange(R)] = 0
        assert np.abs(back - Vz).max() <= bound.max() / 256.0 ** ns


@pytest.mark.parametrize("R,n,k", [(17, 100, 5), (3, 33000, 10), (50, 257, 257), (4, 9, 1)])
def test_topk_bit_exact(dev, R, n, k):
    rng = np.random.default_rng(n)
    S = np.round(rng.random((R, n)), 2)                   # many exact ties
    S[0, :7] = np.nan
    Sd = torch.from_numpy(S).to(dev)
    idx, vals = engine.topk_rows(Sd, k)
    keyed = np.where(np.isnan(S), -np.inf, S)
    want = np.argsort(-keyed, axis=1, kind="stable")[:, :k]
    np.testing.assert_array_equal(idx.cpu().numpy(), want)
    np.testing.assert_array_equal(vals.cpu().numpy(), np.take_along_axis(S, want, axis=1))


def test_topk_matches_oracle(dev):
    rng = np.random.default_rng(11)
    S = rng.random((40, 300))
    idx, vals = engine.topk_rows(torch.from_numpy(S).to(dev), 12)
    oi, ov = orc.topk(S, 12)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi)
    np.testing.assert_array_equal(vals.cpu().numpy(), ov)


# ------------------------------------------------------------------------------- tcgen05 path
def _i8_available():
    return torch.cuda.is_available() and bool(_lib.load().srk_i8_supported())


needs_i8 = pytest.mark.skipif(not _i8_available(), reason="tcgen05 kind::i8 needs sm_100")


def _pad_u8(a, ld):
    out = np.zeros((a.shape[0], ld), dtype=np.uint8)
    out[:, : a.shape[1]] = a
    return out


@needs_i8
@pytest.mark.parametrize("M,K", [(10, 10), (128, 128), (129, 257), (300, 1000), (1000, 130), (700, 4100)])
def test_i8_counts_exact(dev, M, K):
    rng = np.random.default_rng(M + K)
    A = (rng.random((M, K)) < 0.3).astype(np.uint8)
    A[M // 2] = 1                                          # a full row: counts up to K
    op = graph.operator_from_edges(*np.nonzero(A), M, K)
    dop = engine.DeviceOperator(op, dev)
    cnt = dop.evidence_counts("i8")
    torch.cuda.synchronize()
    want = np.minimum(A.astype(np.int64) @ A.astype(np.int64).T, 255).astype(np.uint8)
    np.testing.assert_array_equal(cnt[:, :M].cpu().numpy(), want)


def _planes_of(q, ns, ld):
    R, K = q.shape
    out = np.zeros((ns, R, ld), dtype=np.uint8)
    for s in range(ns):
        out[s, :, :K] = (q >> (8 * (ns - 1 - s))) & 0xFF
    return out


@needs_i8
@pytest.mark.parametrize("ns", [2, 3, 4])
@pytest.mark.parametrize("R,N,K", [(10, 10, 10), (128, 160, 128), (200, 333, 515), (513, 170, 129)])
def test_i8_mid_against_integer_matmul(dev, ns, R, N, K):
    """MID: planes in -> exact integer product -> (+ unit diagonal) -> re-quantised, transposed."""
    rng = np.random.default_rng(ns * 100 + R)
    qmax = 256 ** ns
    q = rng.integers(0, qmax, (R, K), dtype=np.int64)
    A = (rng.random((N, K)) < 0.2).astype(np.uint8)
    ldk, ldr = engine._round_up(K, 128), engine._round_up(R, 128)
    planes = torch.from_numpy(_planes_of(q, ns, ldk)).to(dev)
    a8 = torch.from_numpy(_pad_u8(A, ldk)).to(dev)
    in_vec = torch.from_numpy(rng.random(R) + 0.5).to(dev)
    out_vec = torch.from_numpy(rng.integers(1, 50, N).astype(np.float64)).to(dev)
    out = torch.zeros((ns, N, ldr), dtype=torch.uint8, device=dev)
    a = _lib.I8Args()
    a.mode, a.ns, a.R, a.N, a.K = _lib.SRK_I8_MID, ns, R, N, K
    a.in_planes, a.ld_in, a.in_plane_stride = planes.data_ptr(), ldk, R * ldk
    a.in_rowbound = _lib.RowBound.of(in_vec.data_ptr(), 0.9, 0.0)
    a.A8, a.lda = a8.data_ptr(), ldk
    unit = 1 if R <= K else 0
    a.diag_offset, a.unit_diag = 0, unit
    a.out_planes, a.ld_outp, a.out_plane_stride = out.data_ptr(), ldr, N * ldr
    a.out_rowbound = _lib.RowBound.of(out_vec.data_ptr(), 1.5, 1.0)
    _lib.check(_lib.load().srk_i8_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    D = q @ A.astype(np.int64).T                                        # exact
    inb = in_vec.cpu().numpy() * 0.9
    outb = out_vec.cpu().numpy() * 1.5 + 1.0
    U = D.astype(np.float64) * (inb / qmax)[:, None]
    if unit:
        U = U + A[:, :R].T.astype(np.float64)
    want = np.clip(np.rint(U * (qmax / outb)[None, :]), 0, qmax - 1).astype(np.int64).T      # [N, R]
    p = out.cpu().numpy().astype(np.int64)
    got = sum(p[s] << (8 * (ns - 1 - s)) for s in range(ns))[:, :R]
    assert np.abs(got - want).max() <= 1, np.abs(got - want).max()
    assert (got != want).mean() < 1e-3


@needs_i8
@pytest.mark.parametrize("ns", [2, 3, 4])
@pytest.mark.parametrize("R,K,with_extras", [(10, 10, False), (160, 128, True), (333, 515, True), (515, 129, False)])
def test_i8_final_against_integer_matmul(dev, ns, R, K, with_extras):
    rng = np.random.default_rng(ns * 10 + R)
    N = R
    qmax = 256 ** ns
    q = rng.integers(0, qmax, (R, K), dtype=np.int64)
    A = (rng.random((N, K)) < 0.2).astype(np.uint8)
    ldk, ldn = engine._round_up(K, 128), engine._round_up(N, 16)
    ldp = engine._round_up(N, 128)
    planes = torch.from_numpy(_planes_of(q, ns, ldk)).to(dev)
    a8 = torch.from_numpy(_pad_u8(A, ldk)).to(dev)
    in_vec = torch.from_numpy(rng.random(R) + 0.5).to(dev)
    g = rng.random(N) * 0.01
    gd = torch.from_numpy(g).to(dev)
    S_old = rng.random((R, N))
    S = torch.zeros((R, ldn), dtype=torch.float64, device=dev)
    S[:, :N] = torch.from_numpy(S_old)
    out_planes = torch.zeros((ns, R, ldp), dtype=torch.uint8, device=dev)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    a = _lib.I8Args()
    a.mode, a.ns, a.R, a.N, a.K = _lib.SRK_I8_FINAL, ns, R, N, K
    a.in_planes, a.ld_in, a.in_plane_stride = planes.data_ptr(), ldk, R * ldk
    a.in_rowbound = _lib.RowBound.of(in_vec.data_ptr(), 2.0, 1.0)
    a.A8, a.lda = a8.data_ptr(), ldk
    a.g_row = a.g_col = gd.data_ptr()
    a.out_f64, a.ld_out = S.data_ptr(), ldn
    a.out_planes, a.ld_outp, a.out_plane_stride = out_planes.data_ptr(), ldp, R * ldp
    e = a.epi
    e.coef = 0.8
    e.s_old, e.ld_s_old = S.data_ptr(), ldn
    e.maxdiff, e.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    cnt = prior = None
    if with_extras:
        cnt = rng.integers(0, 60, (R, N)).astype(np.uint8)
        ev = torch.zeros((R, ldn), dtype=torch.uint8, device=dev)
        ev[:, :N] = torch.from_numpy(cnt)
        prior = rng.random((R, N))
        pr = torch.from_numpy(prior).to(dev)
        e.evidence, e.ld_evidence = ev.data_ptr(), ldn
        e.prior, e.ld_prior, e.lambda_ = pr.data_ptr(), N, 0.25
    D = (q @ A.astype(np.int64).T).astype(np.float64)
    inb = in_vec.cpu().numpy() * 2.0 + 1.0
    want = D * (inb / qmax)[:, None] * g[:, None] * g[None, :] * 0.8
    if with_extras:
        want = (1 - 0.25) * want * (1 - 0.5 ** cnt.astype(np.int64)) + 0.25 * prior
    np.fill_diagonal(want, 1.0)
    off = want.copy()
    np.fill_diagonal(off, 0.0)
    bound = off.max() * 1.001
    a.out_rowbound = _lib.RowBound.of(None, 0.0, bound)
    _lib.check(_lib.load().srk_i8_half(C.byref(a), engine._stream()))
    torch.cuda.synchronize()
    got = S[:, :N].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    md, mo = scal.tolist()
    assert md == np.abs(got - S_old).max()
    goff = got.copy()
    np.fill_diagonal(goff, 0.0)
    assert mo == goff.max()
    p = out_planes.cpu().numpy().astype(np.int64)
    qq = sum(p[s] << (8 * (ns - 1 - s)) for s in range(ns))[:, :N]
    wq = np.clip(np.rint(goff * (qmax / bound)), 0, qmax - 1).astype(np.int64)
    assert np.abs(qq - wq).max() <= 1
    assert not np.diag(qq).any()
