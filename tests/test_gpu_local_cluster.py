"""The multi-GPU algorithms on ONE GPU: P logical ranks (threads) of a ``dist.LocalCluster`` run the
unmodified row-sharded solvers on the real kernels -- peer stores of the U blocks, mirrored S blocks
and row-maximum keys land in the other logical ranks' buffers on the same device, the staged variant
moves them with the emulated all-to-all (SURVEY.md section 4 iv).  Runs on the driver's single-GPU
box, so the sharded code paths are exercised there too."""
import numpy as np
import pytest
import torch

from oracle import simrank_oracle as orc
from simrank_b200 import dist as sdist
from simrank_b200 import synth
from SimRank import SimRank as M

pytestmark = pytest.mark.gpu

TOL = {"i8": 1e-6, "csr": 1e-12, "csr16": 1e-6}


def _fit_sharded(world, make, fit_kw, exchange=None, monkeypatch=None):
    if exchange and monkeypatch is not None:
        monkeypatch.setenv("SIMRANK_B200_EXCHANGE", exchange)
    cluster = sdist.LocalCluster(world)

    def body(group):
        obj = make(group)
        out = obj.fit(**fit_kw)
        torch.cuda.synchronize()
        return out, obj.fit_info_

    return cluster.run(body)


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("mode,exchange", [("i8", "auto"), ("i8", "staged"), ("csr", "auto"), ("csr16", "auto")])
def test_directed_logical_shards_match_oracle(world, mode, exchange, monkeypatch):
    df = synth.directed_frame(1000, 20000, 0.8, 21)            # 1000 rows: uneven blocks for every world
    nodes, So, ko, co = orc.fit_directed(df, iterations=50, eps=1e-4)
    res = _fit_sharded(world, lambda g: M.SimRank(mode=mode, sharded=g),
                       dict(data=df, iterations=50, eps=1e-4, verbose=False), exchange, monkeypatch)
    for S, info in res:                                        # gather="all": every rank returns the whole matrix
        assert list(S.index) == nodes and info.mode == mode
        assert (info.applied, info.converged) == (ko, co)      # the convergence decision is global
        assert np.abs(S.to_numpy() - So).max() <= TOL[mode]
    for S, _ in res[1:]:
        np.testing.assert_array_equal(S.to_numpy(), res[0][0].to_numpy())


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("mode", ["i8", "csr", "csr16"])
def test_bipartite_pp_logical_shards_match_oracle(world, mode):
    """BASELINE cfg5 at 1/32 scale (n1 != n2: Evidence_N2 for group 2), local row blocks."""
    df = synth.config_frame("cfg5", scale=1 / 32)
    l1, l2, S1o, S2o, _, _ = orc.fit_bipartite(df, kind="simrank_pp", weighted=True, iterations=3, eps=0.0)
    res = _fit_sharded(world, lambda g: M.BipartitleSimRankPP(mode=mode, sharded=g, gather="local"),
                       dict(data=df, weighted=True, iterations=3, eps=0.0, verbose=False))
    pos1, pos2 = {lab: i for i, lab in enumerate(l1)}, {lab: i for i, lab in enumerate(l2)}
    seen1 = seen2 = 0
    for (S1, S2), info in res:                                 # gather="local": every rank returns its row blocks
        r1, r2 = [pos1[x] for x in S1.index], [pos2[x] for x in S2.index]
        assert list(S1.columns) == l1 and list(S2.columns) == l2
        if r1:
            assert np.abs(S1.to_numpy() - S1o[r1]).max() <= TOL[mode]
        if r2:
            assert np.abs(S2.to_numpy() - S2o[r2]).max() <= TOL[mode]
        seen1, seen2 = seen1 + len(r1), seen2 + len(r2)
    assert (seen1, seen2) == (len(l1), len(l2))                # the blocks tile both matrices


def test_bipartite_csr16_shards_with_split_hub_rows_are_bit_identical(monkeypatch):
    """The popular items of cfg5 have thousands of raters: their neighbour lists are pre-summed in pieces
    (engine.ListSplit).  Integer sums in another order: the sharded result does not change by a bit."""
    df = synth.config_frame("cfg5", scale=1 / 32)
    out = []
    for min_deg, via_accum, first in (("0", "0", "0"), ("64", "0", "0"), ("64", "1", "0"), ("64", "1", "1")):
        monkeypatch.setenv("SRK_SPLIT_MIN", min_deg)
        monkeypatch.setenv("SRK_FINAL_VIA_ACCUM", via_accum)       # second half as ACCUM + FINISH
        monkeypatch.setenv("SRK_FIRST_VIA_ACCUM", first)           # first half as ACCUM + FINISH_FIRST
        monkeypatch.setenv("SRK_SPLIT_PIECE", "48")
        monkeypatch.setenv("SRK_SPLIT_RANGE_MB", "1")
        res = _fit_sharded(3, lambda g: M.BipartitleSimRankPP(mode="csr16", sharded=g, gather="local"),
                           dict(data=df, weighted=True, iterations=3, eps=0.0, verbose=False))
        out.append([(S1.to_numpy(), S2.to_numpy()) for (S1, S2), _ in res])
    for other in out[1:]:
        for (a1, a2), (b1, b2) in zip(out[0], other):
            np.testing.assert_array_equal(a1, b1)
            np.testing.assert_array_equal(a2, b2)


def test_local_cluster_propagates_errors():
    cluster = sdist.LocalCluster(2)

    def body(group):
        if group.rank == 1:
            raise ValueError("boom")
        group.barrier()

    with pytest.raises(ValueError, match="boom"):
        cluster.run(body)
