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
def test_evidence_counts_tensor_core_uint8(dev, M, K):
    """The uint8 evidence counts of SimRank.py:315 through srk_x2_half COUNTS (counts_bits = 8)."""
    rng = np.random.default_rng(M + K)
    A = (rng.random((M, K)) < 0.3).astype(np.uint8)
    A[M // 2] = 1                                          # a full row: counts up to K
    op = graph.operator_from_edges(*np.nonzero(A), M, K)
    dop = engine.DeviceOperator(op, dev)
    cnt = dop.evidence_counts("i8")
    torch.cuda.synchronize()
    want = np.minimum(A.astype(np.int64) @ A.astype(np.int64).T, 255).astype(np.uint8)
    np.testing.assert_array_equal(cnt[:, :M].cpu().numpy(), want)


# ------------------------------------------------------------------------------- edge list -> CSR on the device
def _edges_to_csr(dev, rows, cols, M, K):
    lib = _lib.load()
    m = len(rows)
    r = torch.from_numpy(np.asarray(rows, dtype=np.int32)).to(dev)
    c = torch.from_numpy(np.asarray(cols, dtype=np.int32)).to(dev)
    indptr = torch.full((M + 1,), -1, dtype=torch.int64, device=dev)
    indices = torch.full((max(m, 1),), -1, dtype=torch.int32, device=dev)
    status = torch.full((1,), 99, dtype=torch.int32, device=dev)
    nbytes = int(lib.srk_edges_to_csr_workspace(m, M))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.srk_edges_to_csr(engine._ptr(r), engine._ptr(c), m, M, K, engine._ptr(indptr), engine._ptr(indices),
                                    engine._ptr(status), engine._ptr(ws), nbytes, engine._stream()))
    torch.cuda.synchronize()
    return indptr.cpu().numpy(), indices.cpu().numpy()[:m], int(status.item())


@pytest.mark.parametrize("M,K,m", [(1, 1, 1), (50, 70, 400), (5000, 3000, 60000), (300, 200000, 50000),
                                   (20000, 64, 100000)])
def test_edges_to_csr_matches_host_build(dev, M, K, m):
    rng = np.random.default_rng(M + K)
    keys = rng.choice(M * K, size=min(m, M * K), replace=False)         # unique pairs, arbitrary order
    rows, cols = keys // K, keys % K
    if M > 10:
        rows[rows == 3] = 4                                                # an empty row ...
        keep = np.unique(rows * K + cols, return_index=True)[1]            # ... without creating duplicates
        rows, cols = rows[keep], cols[keep]
        perm = rng.permutation(rows.size)
        rows, cols = rows[perm], cols[perm]
    indptr, indices, status = _edges_to_csr(dev, rows, cols, M, K)
    want_ptr, want_idx = graph._csr(rows.astype(np.int64), cols.astype(np.int64), M, K)
    assert status == 0
    np.testing.assert_array_equal(indptr, want_ptr)
    np.testing.assert_array_equal(indices, want_idx)


def test_edges_to_csr_flags_duplicates_and_bad_indices(dev):
    rows, cols = np.array([0, 1, 1, 2, 1]), np.array([1, 2, 0, 2, 2])    # (1, 2) twice
    assert _edges_to_csr(dev, rows, cols, 3, 3)[2] & 1
    assert _edges_to_csr(dev, np.array([0, 5]), np.array([1, 1]), 3, 3)[2] & 2
    assert _edges_to_csr(dev, np.array([0, 1]), np.array([1, -1]), 3, 3)[2] & 2
    indptr, _, status = _edges_to_csr(dev, np.array([], dtype=np.int64), np.array([], dtype=np.int64), 4, 4)
    assert status == 0 and indptr.tolist() == [0, 0, 0, 0, 0]


def test_device_csr_feeds_the_drop_in_classes(dev):
    """engine.device_csr is what drivers.build_* use above DEVICE_CSR_MIN_EDGES: same operator as the
    host build, the pivot's ValueError on duplicate pairs, device arrays reused by DeviceOperator."""
    import pandas as pd

    from simrank_b200 import drivers, synth
    df = synth.directed_frame(3000, 60000, 0.8, 5)
    _, nodes_h, op_h = graph.build_directed(df, False, "from", "to", "weight")
    _, nodes_d, op_d = drivers.build_directed(df, False, "from", "to", "weight")
    assert nodes_h == nodes_d and op_d.dev_csr is not None and op_h.dev_csr is None
    np.testing.assert_array_equal(op_d.indptr, op_h.indptr)
    np.testing.assert_array_equal(op_d.indices, op_h.indices)
    dop = engine.DeviceOperator(op_d, dev)
    assert dop.indptr.data_ptr() == op_d.dev_csr[0].data_ptr()
    dup = pd.concat([df, df.iloc[:1]], ignore_index=True)
    with pytest.raises(ValueError, match="duplicate entries"):
        drivers.build_directed(dup, False, "from", "to", "weight")
    bi = synth.bipartite_frame(700, 300, 40000, 1.0, 23)
    out_h = graph.build_bipartite(bi, True, "user", "item", "weight")
    out_d = drivers.build_bipartite(bi, True, "user", "item", "weight")
    for a, b in zip(out_h[4:], out_d[4:]):
        np.testing.assert_array_equal(a.indptr, b.indptr)
        np.testing.assert_array_equal(a.indices, b.indices)
        np.testing.assert_array_equal(a.g, b.g)
