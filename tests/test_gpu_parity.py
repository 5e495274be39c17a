"""Class-level parity of the drop-in ``SimRank`` package against the oracle and the golden
fixtures (run on the B200 box: ``pytest -m gpu``).

Tolerances: the float64 CSR path differs from numpy only by summation order (1e-12); the
tcgen05 fixed-point path must stay within the north-star bound of max-abs 1e-6 after K
iterations (it lands around 1e-8)."""
import io
from contextlib import redirect_stdout

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import (load_notebook, load_ref_case, load_ref_index, notebook_bipartite_df,
                      notebook_directed_df)
from oracle import simrank_oracle as orc
from simrank_b200 import _lib, synth
from SimRank import SimRank as M

pytestmark = pytest.mark.gpu

TOL = {"csr": 1e-12, "i8": 1e-6, "csr16": 1e-6}
DIRECTED = dict(from_node_column="ORIGIN_AIRPORT_ID", to_node_column="DEST_AIRPORT_ID", weight_column="flights")
BIPART = dict(node_group1_column="userId", node_group2_column="movieId", weight_column="rating")


def _modes():
    ok = torch.cuda.is_available() and bool(_lib.load().srk_i8_supported())
    return ["csr", "i8", "csr16"] if ok else ["csr"]


MODES = _modes()


def _aligned(frame, labels):
    return frame.loc[labels, labels].to_numpy()


# ------------------------------------------------------------------------------- notebook goldens
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,cls,weighted,rtol,atol", [
    ("B1_directed_unweighted", "SimRank", False, 0, 5.1e-7),
    ("B2_directed_weighted", "SimRank", True, 4e-7, 1e-12),
    ("B3_directed_pp_weighted", "SimRankPP", True, 4e-7, 1e-12),
])
def test_notebook_directed(mode, name, cls, weighted, rtol, atol):
    ent = load_notebook()[name]
    buf = io.StringIO()
    with redirect_stdout(buf):
        S = getattr(M, cls)(mode=mode).fit(notebook_directed_df(), weighted=weighted, **DIRECTED)
    t = ent["tables"][0]
    got = S.loc[t["rows"], t["cols"]].to_numpy()
    np.testing.assert_allclose(got, np.array(t["values"]), rtol=rtol, atol=atol)
    assert f"Converged at iteration {ent['converged_at']}" in buf.getvalue()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,cls,weighted", [
    ("B4_bipartite_unweighted", "BipartiteSimRank", False),
    ("B5_bipartite_weighted", "BipartiteSimRank", True),
    ("B6_bipartite_pp_unweighted", "BipartiteSimRankPP", False),
    ("B7_bipartite_pp_weighted", "BipartiteSimRankPP", True),
])
def test_notebook_bipartite(mode, name, cls, weighted):
    ent = load_notebook()[name]
    buf = io.StringIO()
    with redirect_stdout(buf):
        S1, S2 = getattr(M, cls)(mode=mode).fit(notebook_bipartite_df(), weighted=weighted, **BIPART)
    np.testing.assert_allclose(S1.to_numpy(), np.array(ent["tables"][0]["values"]), rtol=0, atol=5.1e-7)
    np.testing.assert_allclose(S2.to_numpy(), np.array(ent["tables"][1]["values"]), rtol=0, atol=5.1e-7)
    assert f"Converged at iteration {ent['converged_at']}" in buf.getvalue()


# ------------------------------------------------------------------------------- reference-loop fixtures
_CASES, _ = load_ref_index()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", sorted(_CASES))
def test_reference_loop_fixtures(mode, name):
    meta, df, arr = load_ref_case(name)
    kw = dict(meta["kwargs"])
    cls = getattr(M, meta["class"])
    if meta["family"] == "directed":
        labels = arr["labels"].tolist()
        args = ()
        if meta["class"] == "AprioriSimRank" and mode == "csr16":
            with pytest.raises(ValueError, match="prior"):      # the mirrored second half takes no prior
                cls(mode=mode).fit(df, arr["prior"], verbose=False, **kw)
            return
        if meta["class"] == "AprioriSimRank":
            nodes = list(set(df["from"].unique()) | set(df["to"].unique()))
            pos = {n: i for i, n in enumerate(labels)}
            p = [pos[x] for x in nodes]
            args = (arr["prior"][np.ix_(p, p)],)
        obj = cls(mode=mode)
        S = obj.fit(df, *args, verbose=False, **kw)
        np.testing.assert_allclose(_aligned(S, labels), arr["S"], rtol=0, atol=TOL[mode])
        assert np.all(np.diag(S.to_numpy()) == 1.0)
    else:
        args = ()
        if meta["class"] == "BipartitleAprioriSimRank":
            if mode == "csr16":
                with pytest.raises(ValueError, match="prior"):
                    cls(mode=mode).fit(df, arr["prior1"], arr["prior2"], verbose=False, **kw)
                return
            args = (arr["prior1"], arr["prior2"])          # positional in sorted-label order (SimRank.py:488,491)
        obj = cls(mode=mode)
        S1, S2 = obj.fit(df, *args, verbose=False, **kw)
        assert list(S1.index) == arr["sorted1"].tolist() and list(S2.index) == arr["sorted2"].tolist()
        np.testing.assert_allclose(S1.to_numpy(), arr["S1"], rtol=0, atol=TOL[mode])
        np.testing.assert_allclose(S2.to_numpy(), arr["S2"], rtol=0, atol=TOL[mode])
        ref = cls(mode=mode, label_order="reference").fit(df, *args, verbose=False, **kw)
        assert set(ref[0].index) == set(S1.index)
        np.testing.assert_array_equal(ref[0].to_numpy(), S1.to_numpy())
    info = obj.fit_info_
    if mode == "csr" or kw.get("eps", 1e-4) > 0:
        assert (info.applied if info.converged else -1) == meta["converged_at"]
    else:
        # eps == 0 stops only at a bit-exact fixed point of the arithmetic in use; the fixed-point
        # path reaches its own (within TOL of the float64 one, checked above) a little earlier
        assert info.applied <= kw["iterations"]


# ------------------------------------------------------------------------------- BASELINE configs
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("weighted", [False, True])
def test_cfg1_airport_graph(mode, weighted):
    df = synth.config_frame("cfg1")
    for eps, iters in ((1e-4, 10), (0.0, 10)):
        nodes, So, ko, co = orc.fit_directed(df, weighted=weighted, iterations=iters, eps=eps)
        obj = M.SimRank(mode=mode)
        S = obj.fit(df, weighted=weighted, iterations=iters, eps=eps, verbose=False)
        assert list(S.index) == nodes
        assert np.abs(S.to_numpy() - So).max() <= TOL[mode]
        if mode == "csr" or eps > 0:                   # eps == 0: see test_reference_loop_fixtures
            assert (obj.fit_info_.applied, obj.fit_info_.converged) == (ko, co)


@pytest.mark.parametrize("mode", MODES)
def test_cfg2_simrank_pp(mode):
    df = synth.config_frame("cfg2")
    nodes, So, ko, co = orc.fit_directed(df, kind="simrank_pp", weighted=True, iterations=10, eps=0.0)
    obj = M.SimRankPP(mode=mode)
    S = obj.fit(df, weighted=True, iterations=10, eps=0.0, verbose=False)
    assert np.abs(S.to_numpy() - So).max() <= TOL[mode]
    # the lazily materialised attributes are the reference's ndarrays
    _, G = orc.directed_graph(df, True)
    np.testing.assert_array_equal(np.asarray(obj.Evidence), orc.evidence(G))
    np.testing.assert_allclose(np.asarray(obj.Weight), orc.weight(G), rtol=1e-15)
    np.testing.assert_allclose(obj.Graph.to_numpy(), G, rtol=1e-15)


@pytest.mark.parametrize("mode", MODES)
def test_cfg3_bipartite(mode):
    df = synth.config_frame("cfg3")
    l1, l2, S1o, S2o, ko, co = orc.fit_bipartite(df, weighted=True, iterations=10, eps=0.0)
    obj = M.BipartitleSimRank(mode=mode)                  # README spelling alias
    S1, S2 = obj.fit(df, weighted=True, iterations=10, eps=0.0, verbose=False)
    assert list(S1.index) == l1 and list(S2.index) == l2
    assert np.abs(S1.to_numpy() - S1o).max() <= TOL[mode]
    assert np.abs(S2.to_numpy() - S2o).max() <= TOL[mode]
    df_u = synth.config_frame("cfg3")
    l1, l2, S1o, S2o, ko, co = orc.fit_bipartite(df_u, weighted=False, iterations=6, eps=0.0)
    S1, S2 = M.BipartiteSimRank(mode=mode).fit(df_u, weighted=False, iterations=6, eps=0.0, verbose=False)
    assert np.abs(S1.to_numpy() - S1o).max() <= TOL[mode] and np.abs(S2.to_numpy() - S2o).max() <= TOL[mode]


@pytest.mark.parametrize("mode", MODES)
def test_cfg5_scaled_bipartite_pp_rectangular(mode):
    """1/64-scale cfg5: n1 != n2, where the reference raises (SimRank.py:423) and Evidence_N2 is
    used for the group-2 update (oracle.pp_group2_evidence)."""
    df = synth.config_frame("cfg5", scale=1 / 64)
    l1, l2, S1o, S2o, ko, co = orc.fit_bipartite(df, kind="simrank_pp", weighted=True, iterations=3, eps=0.0)
    S1, S2 = M.BipartitleSimRankPP(mode=mode).fit(df, weighted=True, iterations=3, eps=0.0, verbose=False)
    assert S1.shape[0] != S2.shape[0]
    assert np.abs(S1.to_numpy() - S1o).max() <= TOL[mode]
    assert np.abs(S2.to_numpy() - S2o).max() <= TOL[mode]


@pytest.mark.parametrize("mode", MODES)
def test_cfg4_scaled_dense_regime(mode):
    """1/8-scale cfg4 (n=4096, mean in-degree 64): same degree structure as the headline run."""
    df = synth.directed_frame(4096, 4096 * 64, 0.5, 4)
    nodes, So, ko, co = orc.fit_directed(df, iterations=5, eps=0.0)
    S = M.SimRank(mode=mode).fit(df, iterations=5, eps=0.0, verbose=False)
    err = np.abs(S.to_numpy() - So).max()
    assert err <= TOL[mode], err
    Sv = S.to_numpy()
    assert np.all(np.diag(Sv) == 1.0) and Sv.min() >= 0.0 and Sv.max() <= 1.0
    assert np.abs(Sv - Sv.T).max() == 0.0               # every mode mirrors the pairs it computes once


# ------------------------------------------------------------------------------- behaviour
def test_verbose_output_matches_reference_format():
    df = notebook_directed_df()
    buf = io.StringIO()
    with redirect_stdout(buf):
        M.SimRank(mode="csr").fit(df, iterations=4, eps=1e-4, **DIRECTED)
    bar = lambda f: "\rPercent: [" + "#" * int(round(30 * f)) + "-" * (30 - int(round(30 * f))) + f"] {round(f * 100, 1)}% "
    assert buf.getvalue() == "Start iterating...\n" + "".join(bar(i / 4) for i in range(4))   # no completion line
    buf = io.StringIO()
    with redirect_stdout(buf):
        M.SimRankPP(mode="csr").fit(df, weighted=True, **DIRECTED)
    out = buf.getvalue()
    assert out.startswith("Initializing Weight matrix...\nFinished in ")
    assert "s!\nInitializing Evidence matrix...\nFinished in " in out
    assert out.endswith("s!\nStart iterating...\n" + bar(0.0) +
                        "\rPercent: [" + "#" * 30 + "] 100% Complete! \n\rConverged at iteration 1")


def test_edge_cases_identity_results():
    df = pd.DataFrame({"from": [1, 2, 3], "to": [2, 3, 1]})
    for kw in (dict(iterations=0), dict(eps=1.0)):
        obj = M.SimRank(mode="csr")
        S = obj.fit(df, verbose=False, **kw)
        assert np.array_equal(S.to_numpy(), np.eye(3)) and obj.fit_info_.applied == 0
    assert obj.Nodes == {1, 2, 3} and obj.Graph.shape == (3, 3)
    before = df.copy()
    M.SimRank(mode="csr").fit(df, verbose=False)
    pd.testing.assert_frame_equal(df, before)              # caller's frame is not mutated


def test_nodes_without_in_edges_and_negative_weights():
    df = pd.DataFrame({"from": [1, 1, 2, 5, 5], "to": [2, 3, 3, 2, 3], "weight": [1.0, 2.0, -5.0, 0.5, 1.0]})
    for weighted in (False, True):
        nodes, So, ko, co = orc.fit_directed(df, weighted=weighted, iterations=6, eps=0.0)
        S = M.SimRank().fit(df, weighted=weighted, iterations=6, eps=0.0, verbose=False)     # auto mode
        np.testing.assert_allclose(_aligned(S, nodes), So, rtol=0, atol=1e-12)


@pytest.mark.parametrize("mode", MODES)
def test_top_k_bit_exact_on_result(mode):
    df = synth.config_frame("cfg1")
    obj = M.SimRank(mode=mode)
    S = obj.fit(df, iterations=10, eps=0.0, verbose=False)
    lab, val = obj.top_k(10)
    oi, ov = orc.topk(S.to_numpy(), 10)                    # oracle top-k of the SAME matrix: bit-exact
    np.testing.assert_array_equal(lab.to_numpy(), np.asarray(S.index, dtype=object)[oi])
    np.testing.assert_array_equal(val.to_numpy(), ov)
    assert list(lab.iloc[:, 0]) == list(S.index)           # every node is its own nearest neighbour


@pytest.mark.parametrize("mode", MODES)
def test_top_k_matches_the_oracle_where_rank_gaps_exceed_the_tolerance(mode):
    """north_star: node-index / top-k outputs bit-exact against the REFERENCE.  The fixed-point paths
    may deviate by up to 1e-6 per value, so an index can only be demanded where the oracle's
    neighbouring values are further apart than 2e-6; there it must be identical."""
    df = synth.directed_frame(1500, 30000, 1.0, 22, weights="lognormal")
    nodes, So, _, _ = orc.fit_directed(df, weighted=True, iterations=6, eps=0.0)
    obj = M.SimRank(mode=mode)
    S = obj.fit(df, weighted=True, iterations=6, eps=0.0, verbose=False)
    assert list(S.index) == nodes
    k = 10
    lab, _ = obj.top_k(k)
    oi, ov = orc.topk(So, k + 1)
    gap = ov[:, :-1] - ov[:, 1:]                            # gap[p] separates rank p from rank p + 1
    clear = gap[:, :k] > 2e-6
    clear[:, 1:] &= gap[:, : k - 1] > 2e-6                  # ... and from rank p - 1
    assert clear.mean() > 0.5                               # the test has teeth
    want = np.asarray(nodes, dtype=object)[oi[:, :k]]
    assert np.array_equal(lab.to_numpy()[clear], want[clear])


def test_auto_mode_picks_the_path_by_size_and_density():
    """engine.choose_mode: dense-ish graphs -> tensor-core chain, sparse ones -> fixed-point gather,
    small graphs and graphs the planes cannot hold -> float64 (DESIGN.md "mode selection")."""
    cases = [(synth.directed_frame(4096, 4096 * 64, 0.5, 4), {}, "i8"),          # 1.6 % dense
             (synth.directed_frame(4096, 4096 * 8, 0.5, 5), {}, "csr16"),        # 0.2 % dense
             (synth.directed_frame(700, 9000, 1.0, 7), {}, "csr")]               # small: exact arithmetic is cheap
    dfw = synth.directed_frame(2000, 40000, 1.0, 31, weights="lognormal")
    dfw["weight"] = dfw["weight"] - 0.5                                          # negative weight sums
    cases.append((dfw, dict(weighted=True), "csr"))
    for df, kw, want in cases:
        obj = M.SimRank()
        S = obj.fit(df, iterations=3, eps=0.0, verbose=False, **kw)
        assert obj.fit_info_.mode == want, (obj.fit_info_.mode, want)
        nodes, So, _, _ = orc.fit_directed(df, iterations=3, eps=0.0, **kw)
        scale = max(1.0, float(np.abs(So).max()))
        assert np.abs(_aligned(S, nodes) - So).max() / scale <= TOL[want]
