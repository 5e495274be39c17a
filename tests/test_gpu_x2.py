"""Paired-SM tensor-core kernel (srk_x2_half, tcgen05 cta_group::2) and the exact-bound slicer,
checked through the C ABI against exact integer matmuls in numpy (run on the B200 box).

The kernel's integer part is exact, so plane outputs must agree with the numpy restatement up to
one quantisation step on rounding ties and the float64 outputs to 1e-12 relative."""
import ctypes as C

import numpy as np
import pytest
import torch

from simrank_b200 import _lib, engine

pytestmark = pytest.mark.gpu


def _available():
    return torch.cuda.is_available() and bool(_lib.load().srk_i8_supported())


needs_i8 = pytest.mark.skipif(not _available(), reason="tcgen05 kind::i8 needs sm_100")


@pytest.fixture(scope="module")
def dev():
    return engine.require_cuda()


def _pad_u8(a, ld):
    out = np.zeros((a.shape[0], ld), dtype=np.uint8)
    out[:, : a.shape[1]] = a
    return out


def _planes_of(q, ns, ld):
    R, K = q.shape
    out = np.zeros((ns, R, ld), dtype=np.uint8)
    for s in range(ns):
        out[s, :, :K] = (q >> (8 * (ns - 1 - s))) & 0xFF
    return out


def _join(planes, ns):
    p = planes.astype(np.int64)
    return sum(p[s] << (8 * (ns - 1 - s)) for s in range(ns))


def _pow2_exponent(b, ns):
    """f with 2^f the smallest power of two >= b, clamped to f >= 8 ns - 46 (the kernel's rounding
    of the row bounds of U)."""
    m, e = np.frexp(np.asarray(b, dtype=np.float64))          # b = m * 2^e, m in [0.5, 1)
    f = np.where(m == 0.5, e - 1, e)
    return np.maximum(f, 8 * ns - 40).astype(np.int64)


def _exact_matmul(A, q):
    """A (0/1) @ q.T as exact int64: through BLAS in float64 while every sum stays below 2^53."""
    if q.size and float(q.max()) * A.shape[1] < 2.0 ** 52:
        return np.rint(A.astype(np.float64) @ q.T.astype(np.float64)).astype(np.int64)
    return A.astype(np.int64) @ q.T


def _run(a):
    engine.attach_sync_ws(a, torch.device("cuda", torch.cuda.current_device()))   # lockstep scratch (used from K >= 2048)
    _lib.check(_lib.load().srk_x2_half(C.byref(a), engine._stream()), "srk_x2_half")
    torch.cuda.synchronize()


@needs_i8
@pytest.mark.parametrize("M,R,K", [(10, 10, 10), (256, 256, 128), (257, 300, 129), (700, 513, 1000), (130, 1000, 4100)])
def test_x2_counts_exact(dev, M, R, K):
    rng = np.random.default_rng(M + K)
    A = (rng.random((M, K)) < 0.3).astype(np.uint8)
    B = (rng.random((R, K)) < 0.3).astype(np.uint8)
    A[M // 2] = 1
    B[R // 3] = 1                                         # a count equal to K
    ldk = engine._round_up(K, 128)
    a8 = torch.from_numpy(_pad_u8(A, ldk)).to(dev)
    b8 = torch.from_numpy(_pad_u8(B, ldk)).to(dev)
    ldc = engine._round_up(R, 8)
    out = torch.full((M, ldc), 7, dtype=torch.int16, device=dev)
    a = _lib.X2Args()
    a.mode, a.ns, a.M, a.R, a.K = _lib.SRK_X2_COUNTS, 1, M, R, K
    a.A8, a.lda = a8.data_ptr(), ldk
    a.in_planes, a.ld_in, a.in_plane_stride = b8.data_ptr(), ldk, R * ldk
    a.out_counts, a.ld_out_counts = out.data_ptr(), ldc
    _run(a)
    want = np.minimum(A.astype(np.int64) @ B.astype(np.int64).T, 65535)
    got = out.cpu().numpy().view(np.uint16)[:, :R].astype(np.int64)
    np.testing.assert_array_equal(got, want)


@needs_i8
def test_x2_counts_uint32(dev):
    """Counts beyond 65535 need the 32-bit layout (rows sharing more than 65535 neighbours)."""
    M, R, K = 200, 150, 70016
    rng = np.random.default_rng(9)
    A = (rng.random((M, K)) < 0.5).astype(np.uint8)
    B = (rng.random((R, K)) < 0.5).astype(np.uint8)
    A[3] = 1
    B[5] = 1                                              # count 70016 > 65535
    a8, b8 = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    ldc = engine._round_up(R, 8)
    out = torch.full((M, ldc), 7, dtype=torch.int32, device=dev)
    a = _lib.X2Args()
    a.mode, a.ns, a.M, a.R, a.K = _lib.SRK_X2_COUNTS, 1, M, R, K
    a.A8, a.lda = a8.data_ptr(), K
    a.in_planes, a.ld_in, a.in_plane_stride = b8.data_ptr(), K, R * K
    a.out_counts, a.ld_out_counts, a.counts_bits = out.data_ptr(), ldc, 32
    _run(a)
    want = A.astype(np.int64) @ B.astype(np.int64).T
    assert want.max() == K
    np.testing.assert_array_equal(out.cpu().numpy()[:, :R].astype(np.int64), want)


@needs_i8
@pytest.mark.parametrize("ns", [2, 3, 4])
@pytest.mark.parametrize("M,R,K", [(10, 10, 10), (256, 128, 128), (333, 200, 515), (170, 513, 129), (600, 70, 1300)])
def test_x2_mid_against_integer_matmul(dev, ns, M, R, K):
    rng = np.random.default_rng(ns * 100 + M)
    qmax = 256 ** ns
    q = rng.integers(0, qmax, (R, K), dtype=np.int64)
    A = (rng.random((M, K)) < 0.2).astype(np.uint8)
    ldk, ldr = engine._round_up(K, 128), engine._round_up(R, 128)
    planes = torch.from_numpy(_planes_of(q, ns, ldk)).to(dev)
    a8 = torch.from_numpy(_pad_u8(A, ldk)).to(dev)
    in_vec = torch.from_numpy(rng.random(R) + 0.5).to(dev)
    out_vec = torch.from_numpy(rng.integers(1, 50, M).astype(np.float64)).to(dev)
    out = torch.full((ns, M, ldr), 9, dtype=torch.uint8, device=dev)
    a = _lib.X2Args()
    a.mode, a.ns, a.M, a.R, a.K = _lib.SRK_X2_MID, ns, M, R, K
    a.A8, a.lda = a8.data_ptr(), ldk
    a.in_planes, a.ld_in, a.in_plane_stride = planes.data_ptr(), ldk, R * ldk
    a.in_rowbound = _lib.RowBound.of(in_vec.data_ptr(), 0.9, 0.0)
    a.out_planes, a.ld_outp, a.out_plane_stride = out.data_ptr(), ldr, M * ldr
    a.out_rowbound = _lib.RowBound.of(out_vec.data_ptr(), 30.0, 1.0)
    _run(a)
    D = _exact_matmul(A, q).astype(np.float64)                          # [M, R] exact
    inb = in_vec.cpu().numpy() * 0.9
    outb = out_vec.cpu().numpy() * 30.0 + 1.0
    fj = _pow2_exponent(outb, ns)
    want = np.clip(np.rint((D * (2.0 ** -fj)[:, None]) * inb[None, :]), 0, qmax - 1).astype(np.int64)
    got = _join(out.cpu().numpy(), ns)
    np.testing.assert_array_equal(got[:, :R], want)
    pad = got[:, R: engine._round_up(R, 16)]
    assert not pad.any()                                                # chunk padding is written as zeros


def _final_case(rng, ns, M, R, K, layout, extras, dev, bscale=1.0, mirror=False, bits=16):
    qmax = 256 ** ns
    q = rng.integers(0, qmax, (R, K), dtype=np.int64)
    A = (rng.random((M, K)) < 0.2).astype(np.uint8)
    ldk = engine._round_up(K, 128)
    planes = torch.from_numpy(_planes_of(q, ns, ldk)).to(dev)
    a8 = torch.from_numpy(_pad_u8(A, ldk)).to(dev)
    in_vec = torch.from_numpy(rng.random(R) + 0.5).to(dev)
    ga, gv = rng.random(M) * 0.01, rng.random(R) * 0.01
    trans = layout == _lib.SRK_X2_TRANSPOSED
    rows, cols = (R, M) if trans else (M, R)                            # shape of the output matrix
    ld = engine._round_up(cols, 16)
    diag_offset = 3 if trans else 0
    S_old = rng.random((rows, cols))
    cnt = rng.integers(0, 70, (rows, cols)).astype(np.uint16 if bits == 16 else np.uint32)
    cnt[0, : min(cols, 5)] = [65535 if bits == 16 else 300000, 54, 53, 1, 0][: min(cols, 5)]
    if layout == _lib.SRK_X2_SYMMETRIC:
        S_old = np.triu(S_old) + np.triu(S_old, 1).T
        cnt = np.triu(cnt) + np.triu(cnt, 1).T
    prior = rng.random((rows, cols)) if extras == "prior" else None
    S = torch.zeros((rows, ld), dtype=torch.float64, device=dev)
    S[:, :cols] = torch.from_numpy(S_old)
    cd = torch.zeros((rows, ld), dtype=torch.int16 if bits == 16 else torch.int32, device=dev)
    cd[:, :cols] = torch.from_numpy(cnt.view(np.int16 if bits == 16 else np.int32))
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    a = _lib.X2Args()
    a.mode, a.ns, a.layout, a.M, a.R, a.K = _lib.SRK_X2_FINAL, ns, layout, M, R, K
    a.A8, a.lda = a8.data_ptr(), ldk
    a.in_planes, a.ld_in, a.in_plane_stride = planes.data_ptr(), ldk, R * ldk
    a.in_rowbound = _lib.RowBound.of(in_vec.data_ptr(), 2.0 * bscale, 1.0 * bscale)
    gad, gvd = torch.from_numpy(ga).to(dev), torch.from_numpy(gv).to(dev)
    a.g_a, a.g_v = gad.data_ptr(), gvd.data_ptr()
    a.out_f64, a.ld_out, a.diag_offset = S.data_ptr(), ld, diag_offset
    use_counts = extras in ("counts", "evidence", "prior", "ev8")
    if use_counts:
        a.counts, a.ld_counts, a.add_counts, a.counts_bits = cd.data_ptr(), ld, 1, bits
        a.use_evidence = 1 if extras in ("evidence", "prior") else 0
    e = a.epi
    ev8 = None
    if extras == "ev8":                                                 # separate uint8 evidence counts
        ev8 = rng.integers(0, 70, (rows, cols)).astype(np.uint8)
        if layout == _lib.SRK_X2_SYMMETRIC:
            ev8 = np.triu(ev8) + np.triu(ev8, 1).T
        evd = torch.zeros((rows, ld), dtype=torch.uint8, device=dev)
        evd[:, :cols] = torch.from_numpy(ev8)
        e.evidence, e.ld_evidence = evd.data_ptr(), ld
    e.coef = 0.8
    e.s_old, e.ld_s_old = S.data_ptr(), ld
    e.maxdiff, e.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    keep = [planes, a8, in_vec, gad, gvd, cd, locals().get("evd")]
    if prior is not None:
        pr = torch.from_numpy(prior).to(dev)
        keep.append(pr)
        e.prior, e.ld_prior, e.lambda_ = pr.data_ptr(), cols, 0.25
    rk = torch.zeros(rows, dtype=torch.int32, device=dev)               # keys of the row maxima of the output
    a.rowmax_hi = rk.data_ptr()
    mir = mrk = None
    if mirror:                                                          # the block as its owner stores it
        ldm = engine._round_up(R + 6, 2)
        mir = torch.full((M, ldm), -3.0, dtype=torch.float64, device=dev)
        a.mirror_out, a.ld_mirror, a.mirror_col0 = mir.data_ptr(), ldm, 6
        mrk = torch.zeros(M, dtype=torch.int32, device=dev)             # ... and of the mirrored block
        a.mirror_rowmax_hi = mrk.data_ptr()
    _run(a)
    # ---- numpy restatement, in the (j, r) frame of the kernel
    D = _exact_matmul(A, q)                                             # [M, R] exact
    inb = (in_vec.cpu().numpy() * 2.0 + 1.0) * bscale
    s = 8 * ns - _pow2_exponent(inb, ns)                                # bounds of U are powers of two
    dl = np.maximum(-s, 0)
    s = np.maximum(s, 0)
    cnt_jr = (cnt.T if trans else cnt).astype(np.float64)
    T = D << dl[None, :]
    if use_counts:
        T = T + ((cnt.T if trans else cnt).astype(np.int64) << s[None, :])
    val = (T.astype(np.float64) * ((0.8 * gv) * 2.0 ** -s)[None, :]) * ga[:, None]
    if a.use_evidence:
        val = val * (1 - 0.5 ** np.minimum(cnt_jr, 60.0))
    if ev8 is not None:
        val = val * (1 - 0.5 ** (ev8.T if trans else ev8).astype(np.float64))
    if prior is not None:
        val = 0.75 * val + 0.25 * (prior.T if trans else prior)
    jj, rr = np.meshgrid(np.arange(M), np.arange(R), indexing="ij")
    val[jj == rr + diag_offset] = 1.0
    want = val.T if trans else val
    if layout == _lib.SRK_X2_SYMMETRIC:
        want = np.triu(want) + np.triu(want, 1).T
    got = S[:, :cols].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    if mir is not None:
        mh = mir.cpu().numpy()
        exp_m = got.T.copy()
        exp_m[np.arange(M)[:, None] == np.arange(R)[None, :] + diag_offset] = -3.0   # the diagonal has no mirror image
        np.testing.assert_array_equal(mh[:, 6:6 + R], exp_m)             # same values, mirrored placement
        assert np.all(mh[:, :6] == -3.0) and np.all(mh[:, 6 + R:] == -3.0)
    if layout == _lib.SRK_X2_SYMMETRIC:
        assert np.array_equal(got, got.T)
    md, mo = scal.tolist()
    want_md = np.abs(got - S_old).max()                                 # integer subtraction: truncated, < 1 ulp below
    assert want_md * (1 - 2.0 ** -51) <= md <= want_md
    off = got.copy()
    dmask = (jj == rr + diag_offset).T if trans else (jj == rr + diag_offset)
    off[dmask] = 0.0
    assert mo == off.max()
    for keys, m in ((rk, off.max(axis=1)), (mrk, off.max(axis=0))):
        if keys is not None:
            want_key = np.where(m > 0, (m.view(np.int64) >> 32) + 1, 0)
            np.testing.assert_array_equal(keys.cpu().numpy().astype(np.int64), want_key)
    assert not S[:, cols:].cpu().numpy().any()                          # nothing written past the matrix


@needs_i8
@pytest.mark.parametrize("ns", [2, 3, 4])
@pytest.mark.parametrize("M,R,K,extras", [(10, 10, 10, "none"), (256, 128, 128, "counts"), (333, 200, 515, "evidence"),
                                          (170, 513, 129, "prior"), (600, 70, 1300, "ev8")])
def test_x2_final_direct(dev, ns, M, R, K, extras):
    _final_case(np.random.default_rng(ns * 10 + M), ns, M, R, K, _lib.SRK_X2_DIRECT, extras, dev)


@needs_i8
@pytest.mark.parametrize("ns", [2, 3, 4])
@pytest.mark.parametrize("n,K,extras", [(10, 10, "none"), (256, 128, "counts"), (333, 515, "evidence"), (900, 129, "ev8"),
                                        (1500, 260, "evidence")])
def test_x2_final_symmetric(dev, ns, n, K, extras):
    _final_case(np.random.default_rng(ns * 10 + n), ns, n, n, K, _lib.SRK_X2_SYMMETRIC, extras, dev)


@needs_i8
@pytest.mark.parametrize("ns", [2, 3, 4])
@pytest.mark.parametrize("M,R,K,extras", [(10, 7, 10, "none"), (256, 128, 128, "counts"), (515, 200, 333, "evidence"),
                                          (513, 170, 129, "prior"), (300, 90, 200, "ev8")])
def test_x2_final_transposed(dev, ns, M, R, K, extras):
    _final_case(np.random.default_rng(ns * 10 + M), ns, M, R, K, _lib.SRK_X2_TRANSPOSED, extras, dev)


@needs_i8
@pytest.mark.parametrize("ns,bscale", [(2, 3e5), (2, 1e-20), (3, 1e-9), (4, 7.0)])
def test_x2_final_bound_ranges(dev, ns, bscale):
    """Bounds of U above 256^NS (left shift of D) and below the clamp (counts << 46)."""
    _final_case(np.random.default_rng(5), ns, 300, 300, 260, _lib.SRK_X2_SYMMETRIC, "counts", dev, bscale=bscale)
    _final_case(np.random.default_rng(6), ns, 300, 200, 260, _lib.SRK_X2_DIRECT, "evidence", dev, bscale=bscale)


@needs_i8
@pytest.mark.parametrize("layout", [_lib.SRK_X2_DIRECT, _lib.SRK_X2_SYMMETRIC, _lib.SRK_X2_TRANSPOSED])
def test_x2_final_counts_uint32(dev, layout):
    M = R = 400
    _final_case(np.random.default_rng(layout), 2, M, R, 260, layout, "evidence", dev, bits=32)
    _final_case(np.random.default_rng(layout + 5), 3, M, R, 130, layout, "counts", dev, bits=32)


@needs_i8
def test_x2_final_transposed_with_mirror(dev):
    """Row-sharded symmetric update: the local block is stored transposed and its mirror image goes
    to the buffer of the rank that owns row j (here an ordinary device buffer)."""
    for ns, M, R, K in ((2, 512, 256, 384), (3, 700, 300, 200), (2, 130, 70, 129)):
        rng = np.random.default_rng(M)
        # diag_offset of the helper is 3: keep the diagonal out of the picture with M, R as given
        _final_case(rng, ns, M, R, K, _lib.SRK_X2_TRANSPOSED, "counts", dev, mirror=True)


@needs_i8
def test_x2_kblocked_operand(dev):
    """V given as K-blocks (the receive buffer of the row-sharded exchange)."""
    rng = np.random.default_rng(8)
    ns, M, R, nblk, kb = 3, 300, 100, 3, 256
    K = nblk * kb
    qmax = 256 ** ns
    q = rng.integers(0, qmax, (R, K), dtype=np.int64)
    A = (rng.random((M, K)) < 0.2).astype(np.uint8)
    rb = 128                                                            # rows per block in the buffer
    buf = np.zeros((nblk, ns, rb, kb), dtype=np.uint8)
    for b in range(nblk):
        buf[b, :, :R, :] = _planes_of(q[:, b * kb:(b + 1) * kb], ns, kb)
    bufd = torch.from_numpy(buf).to(dev)
    a8 = torch.from_numpy(A).to(dev)
    ldr = engine._round_up(R, 128)
    out = torch.zeros((ns, M, ldr), dtype=torch.uint8, device=dev)
    a = _lib.X2Args()
    a.mode, a.ns, a.M, a.R, a.K = _lib.SRK_X2_MID, ns, M, R, K
    a.A8, a.lda = a8.data_ptr(), K
    a.in_planes, a.ld_in, a.in_plane_stride = bufd.data_ptr(), kb, rb * kb
    a.in_kblock, a.in_kblock_stride = kb, ns * rb * kb
    a.in_rowbound = _lib.RowBound.of(None, 0.0, 1.0)
    a.out_planes, a.ld_outp, a.out_plane_stride = out.data_ptr(), ldr, M * ldr
    a.out_rowbound = _lib.RowBound.of(None, 0.0, float(K))
    _run(a)
    D = _exact_matmul(A, q).astype(np.float64)
    fj = int(_pow2_exponent(float(K), ns))
    want = np.clip(np.rint(D * 2.0 ** -fj), 0, qmax - 1).astype(np.int64)
    got = _join(out.cpu().numpy(), ns)[:, :R]
    np.testing.assert_array_equal(got, want)


@needs_i8
def test_x2_many_tiles_per_pair(dev):
    """More pair tiles than CTA pairs: exercises the persistent loop, both TMEM buffers, the
    shared-memory ring wrap-around and the symmetric tile walk across several bands."""
    _final_case(np.random.default_rng(77), 2, 5000, 5000, 384, _lib.SRK_X2_SYMMETRIC, "counts", dev)
    _final_case(np.random.default_rng(78), 3, 4100, 3000, 256, _lib.SRK_X2_DIRECT, "none", dev)


@needs_i8
def test_x2_lockstep_long_k(dev):
    """K >= 2048 with the scratch attached: the CTA pairs throttle each other through the progress
    counters (four units per tile); results must not depend on it, the launch must not hang, and a
    second launch must find the counters reset."""
    for _ in range(2):
        _final_case(np.random.default_rng(91), 2, 2304, 2304, 8192, _lib.SRK_X2_SYMMETRIC, "counts", dev)
    _final_case(np.random.default_rng(92), 3, 1100, 2100, 4096, _lib.SRK_X2_TRANSPOSED, "evidence", dev, mirror=True)
    test_x2_mid_against_integer_matmul(dev, 2, 2600, 1300, 2176)


@pytest.mark.parametrize("ns", [1, 2, 3, 4])
def test_slice_rows_key(dev, ns):
    """One-pass slicer: bounds come as keys (high word + 1) of the row maxima."""
    rng = np.random.default_rng(40 + ns)
    R, K = 37, 530
    V = rng.random((R, K)) * (rng.random(R) * 3 + 0.01)[:, None]
    V[4] = 0.0
    off = 1
    Vz = V.copy()
    Vz[np.arange(R), np.arange(R) + off] = 0.0
    m = Vz.max(axis=1)
    key = np.where(m > 0, (m.view(np.int64) >> 32) + 1, 0).astype(np.int32)
    mk = (key.astype(np.int64) << 32).view(np.float64)                  # the bound the kernel derives
    assert np.all(mk >= m) and np.all(mk <= m * (1 + 2.0 ** -19) + 1e-300)
    ldp = engine._round_up(K, 128)
    planes = torch.full((ns, R, ldp), 3, dtype=torch.uint8, device=dev)
    bound = torch.zeros(R, dtype=torch.float64, device=dev)
    Vd, kd = torch.from_numpy(V).to(dev), torch.from_numpy(key).to(dev)
    _lib.check(_lib.load().srk_slice_rows_key_f64(engine._ptr(Vd), K, R, K, off, ns, engine._ptr(kd), engine._ptr(planes),
                                                  ldp, R * ldp, engine._ptr(bound), engine._stream()))
    torch.cuda.synchronize()
    qmax = 256.0 ** ns - 1
    np.testing.assert_allclose(bound.cpu().numpy(), mk * ((qmax + 1) / qmax), rtol=1e-15)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = np.where(mk[:, None] > 0, np.rint(Vz * (qmax / np.where(mk > 0, mk, 1.0))[:, None]), 0.0).astype(np.int64)
    got = _join(planes.cpu().numpy(), ns)
    assert np.abs(got[:, :K] - want).max() <= 1
    assert not got[:, K:].any() and not got[4].any()


@pytest.mark.parametrize("ns", [1, 2, 3, 4])
def test_slice_rows_max(dev, ns):
    rng = np.random.default_rng(ns)
    R, K = 45, 1077
    V = rng.random((R, K)) * (rng.random(R) * 3 + 0.01)[:, None]
    V[5] = 0.0                                                          # an all-zero row
    V[7, 3] = -1.0
    V[8, 9] = np.nan
    off = 2
    ldp = engine._round_up(K, 128)
    planes = torch.full((ns, R, ldp), 3, dtype=torch.uint8, device=dev)
    bound = torch.zeros(R, dtype=torch.float64, device=dev)
    Vd = torch.from_numpy(V).to(dev)
    _lib.check(_lib.load().srk_slice_rows_max_f64(engine._ptr(Vd), K, R, K, off, ns, engine._ptr(planes), ldp,
                                                  R * ldp, engine._ptr(bound), engine._stream()))
    torch.cuda.synchronize()
    Vz = np.where(np.isnan(V) | (V < 0), 0.0, V)
    Vz[np.arange(R), np.arange(R) + off] = 0.0
    m = Vz.max(axis=1)
    qmax = 256.0 ** ns - 1
    np.testing.assert_allclose(bound.cpu().numpy(), m * ((qmax + 1) / qmax), rtol=1e-15)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = np.where(m[:, None] > 0, np.rint(Vz * (qmax / m)[:, None]), 0.0).astype(np.int64)
    got = _join(planes.cpu().numpy(), ns)
    assert np.abs(got[:, :K] - want).max() <= 1
    assert got[:, :K].max() == int(qmax)                                # the row maximum maps to the last level
    assert not got[:, K:].any()
    back = got[:, :K] * (bound.cpu().numpy() / (qmax + 1))[:, None]
    assert np.abs(back - Vz).max() <= (m / qmax).max() * 0.5000001
