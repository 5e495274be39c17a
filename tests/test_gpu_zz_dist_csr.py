"""Multi-GPU parity of the row-sharded float64 CSR solver (needs >= 2 GPUs): launches
tests/dist_gpu_worker_csr.py with torchrun, one rank per GPU over NCCL.  Named to run last."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_row_sharded_csr_classes_match_oracle():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tests", "dist_gpu_worker_csr.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_CSR_OK" in res.stdout


def test_csr_half_row_block_calls_single_gpu():
    """The two call shapes of the row-sharded CSR solver on ONE GPU: a first half restricted to the
    graph rows of a destination block that stores into a buffer holding just those columns (shifted
    base), and a second half on a column panel of T with ``diag_offset`` = first row of the block."""
    import ctypes as C

    import numpy as np

    from simrank_b200 import _lib, engine, graph

    dev = engine.require_cuda()
    lib = _lib.load()
    rng = np.random.default_rng(17)
    n, K = 203, 150
    mask = rng.random((n, K)) < 0.1
    rows, cols = np.nonzero(mask)
    op = graph.operator_from_edges(rows, cols, n, K, rng.random(n) * 0.2 - 0.03)     # some negative scales
    dop = engine.DeviceOperator(op, dev)
    G = op.to_dense()
    p = lambda t: engine._ptr(t)                                                      # noqa: E731

    # ---- first half, destination block = graph rows [lo, hi)
    lo, hi, Lx, per = 64, 134, 37, 80
    Xh = rng.random((K, Lx))
    X = torch.from_numpy(Xh).to(dev)
    buf = torch.full((Lx, per), -7.0, dtype=torch.float64, device=dev)
    _lib.check(lib.srk_csr_half_f64(p(dop.indptr), p(dop.indices), p(dop.g), n, lo, hi, p(X), Lx, Lx,
                                    C.c_void_p(buf.data_ptr() - 8 * lo), per, None, engine._stream()))
    torch.cuda.synchronize()
    got = buf.cpu().numpy()
    np.testing.assert_allclose(got[:, : hi - lo], (G @ Xh).T[:, lo:hi], rtol=1e-13, atol=1e-300)
    assert np.all(got[:, hi - lo:] == -7.0)                                          # nothing else touched

    # ---- second half on the column panel T[:, r0 : r0 + L) of a square problem
    op2 = graph.operator_from_edges(*np.nonzero(rng.random((n, n)) < 0.1), n, n, rng.random(n) * 0.2 + 0.01)
    d2 = engine.DeviceOperator(op2, dev)
    G2 = op2.to_dense()
    r0, L = 48, 70
    Th, S_old = rng.random((n, n)), rng.random((n, n))
    cnt = rng.integers(0, 70, (n, n)).astype(np.uint8)
    ld = engine._round_up(n, 16)
    S_blk = torch.zeros((L, ld), dtype=torch.float64, device=dev)
    S_blk[:, :n] = torch.from_numpy(S_old[r0:r0 + L])
    ev = torch.zeros((L, ld), dtype=torch.uint8, device=dev)
    ev[:, :n] = torch.from_numpy(cnt[r0:r0 + L])
    panel = torch.from_numpy(np.ascontiguousarray(Th[:, r0:r0 + L])).to(dev)
    scal = torch.zeros(2, dtype=torch.float64, device=dev)
    e = _lib.Epilogue()
    e.coef = 0.8
    e.evidence, e.ld_evidence = ev.data_ptr(), ld
    e.s_old, e.ld_s_old = S_blk.data_ptr(), ld
    e.maxdiff, e.maxoff = scal.data_ptr(), scal.data_ptr() + 8
    e.diag_offset = r0
    _lib.check(lib.srk_csr_half_f64(p(d2.indptr), p(d2.indices), p(d2.g), n, 0, n, p(panel), L, L, p(S_blk), ld,
                                    C.byref(e), engine._stream()))
    torch.cuda.synchronize()
    want = (1 - 0.5 ** cnt.astype(np.int64)) * 0.8 * (G2 @ Th).T
    np.fill_diagonal(want, 1.0)
    got = S_blk[:, :n].cpu().numpy()
    np.testing.assert_allclose(got, want[r0:r0 + L], rtol=1e-12, atol=1e-300)
    md, mo = scal.tolist()
    assert md == np.abs(got - S_old[r0:r0 + L]).max()
    off = got.copy()
    off[np.arange(L), r0 + np.arange(L)] = 0.0
    assert mo == off.max()


@pytest.mark.parametrize("mode", ["csr", "i8"])
def test_relabelling_the_nodes_permutes_the_result(mode):
    """Permutation equivariance (SURVEY.md section 4 iii): the similarity of two nodes does not depend
    on their labels, on the node order the labels induce, or on where their rows fall in a tile.

    Float64 path: exact up to summation order (1e-13).  Fixed-point path: the planes are cut per row
    and the symmetric FINAL computes each unordered pair once from ONE of the two rows, so which
    rounding a pair sees depends on the node order; the two results can differ by the sum of their
    own guaranteed deviations from the float64 iteration (``fit_info_.error_bound``, the recursion
    e <- kappa e + delta of engine.choose_slices) and by no more, and each is within its bound of the
    oracle."""
    import numpy as np

    from oracle import simrank_oracle as orc
    from simrank_b200 import synth
    from SimRank import SimRank as M

    df = synth.directed_frame(1300, 26000, 0.8, 41)
    labels = np.unique(np.concatenate([df["from"].to_numpy(), df["to"].to_numpy()]))
    relabel = dict(zip(labels.tolist(), (np.random.default_rng(5).permutation(labels.size) * 7 + 3).tolist()))
    df2 = df.assign(**{"from": df["from"].map(relabel), "to": df["to"].map(relabel)})
    df2 = df2.sample(frac=1.0, random_state=3).reset_index(drop=True)               # and another edge order
    o1, o2 = M.SimRank(mode=mode), M.SimRank(mode=mode)
    S1 = o1.fit(df, iterations=4, eps=0.0, verbose=False)
    S2 = o2.fit(df2, iterations=4, eps=0.0, verbose=False)
    mapped = [relabel[x] for x in S1.index]
    diff = np.abs(S1.to_numpy() - S2.loc[mapped, mapped].to_numpy()).max()
    if mode == "csr":
        assert diff <= 1e-13, diff
        return
    b1, b2 = o1.fit_info_.error_bound[0], o2.fit_info_.error_bound[0]
    assert 0.0 < b1 <= 5e-7 and 0.0 < b2 <= 5e-7, (b1, b2)          # engine.ERR_BUDGET
    assert diff <= b1 + b2, (diff, b1, b2)
    nodes, So, _, _ = orc.fit_directed(df, iterations=4, eps=0.0)
    assert list(S1.index) == nodes
    assert np.abs(S1.to_numpy() - So).max() <= b1
    nodes2, So2, _, _ = orc.fit_directed(df2, iterations=4, eps=0.0)
    assert np.abs(S2.loc[nodes2, nodes2].to_numpy() - So2).max() <= b2


def test_empty_edge_list_returns_empty_frames():
    """An edge list without rows: the reference builds 0 x 0 matrices, finds nothing that differs
    (`_converged` of two empty arrays) and returns empty DataFrames at iteration 0."""
    import pandas as pd

    from SimRank import SimRank as M

    df = pd.DataFrame({"from": pd.Series([], dtype="int64"), "to": pd.Series([], dtype="int64"),
                       "weight": pd.Series([], dtype="float64")})
    for cls in (M.SimRank, M.SimRankPP):
        obj = cls()
        S = obj.fit(df, weighted=True, verbose=False)
        assert S.shape == (0, 0) and obj.Nodes == set()
        assert (obj.fit_info_.applied, obj.fit_info_.converged) == (0, True)
    bi = df.rename(columns={"from": "user", "to": "item"})
    obj = M.BipartiteSimRank()
    S1, S2 = obj.fit(bi, verbose=False)
    assert S1.shape == (0, 0) and S2.shape == (0, 0) and obj.fit_info_.converged
