"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the host graph
construction agrees with the oracle's dense restatement, the progress bar is byte-identical,
and the product path refuses to run without a GPU (no CPU fallback)."""
import io
import os
import re
from contextlib import redirect_stdout

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import ROOT, load_ref_case, notebook_bipartite_df, notebook_directed_df
from oracle import simrank_oracle as orc
from simrank_b200 import _lib, graph, synth


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "simrank_b200.h")).read()
    declared = set(re.findall(r"\b(srk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.srk_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.Epilogue) == 11 * 8
    assert C.sizeof(_lib.RowBound) == 24


def test_struct_layouts_match_c_compiler(tmp_path):
    """sizeof/offsetof of every ABI struct as gcc sees the header == the ctypes mirror."""
    import ctypes as C
    import subprocess
    structs = {"srk_epilogue": _lib.Epilogue, "srk_rowbound": _lib.RowBound, "srk_x2_args": _lib.X2Args,
               "srk_csr_args": _lib.CsrArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "simrank_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname.rstrip("_")}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


@pytest.mark.parametrize("weighted", [False, True])
def test_directed_graph_matches_oracle(weighted):
    df = notebook_directed_df()
    kw = dict(from_node_column="ORIGIN_AIRPORT_ID", to_node_column="DEST_AIRPORT_ID", weight_column="flights")
    nodes_o, G = orc.directed_graph(df, weighted, **kw)
    node_set, nodes, op = graph.build_directed(df, weighted, **kw)
    assert nodes == nodes_o and node_set == set(nodes)
    np.testing.assert_array_equal(op.to_dense(), G)
    assert np.all(np.diff(op.indices.astype(np.int64))[np.diff(np.repeat(np.arange(op.M), op.deg)) == 0] > 0)


@pytest.mark.parametrize("name", ["cfg1", "cfg3"])
def test_synthetic_graphs_match_oracle(name):
    df = synth.config_frame(name)
    if name == "cfg1":
        for weighted in (False, True):
            _, G = orc.directed_graph(df, weighted)
            _, _, op = graph.build_directed(df, weighted, "from", "to", "weight")
            np.testing.assert_allclose(op.to_dense(), G, rtol=1e-15, atol=0)
    else:
        l1, l2, G12, G21 = orc.bipartite_graph(df, True)
        s1, s2, o1, o2, op12, op21 = graph.build_bipartite(df, True, "user", "item", "weight")
        assert o1 == l1 and o2 == l2
        np.testing.assert_allclose(op12.to_dense(), G12, rtol=1e-15, atol=0)
        np.testing.assert_allclose(op21.to_dense(), G21, rtol=1e-15, atol=0)


def test_bipartite_graph_matches_oracle_and_reference_fixture():
    meta, df, arr = load_ref_case("bip_sr_w")
    l1, l2, G12, G21 = orc.bipartite_graph(df, True)
    s1, s2, o1, o2, op12, op21 = graph.build_bipartite(df, True, "user", "item", "weight")
    assert o1 == arr["sorted1"].tolist() and o2 == arr["sorted2"].tolist()
    np.testing.assert_array_equal(op12.to_dense(), G12)
    np.testing.assert_array_equal(op21.to_dense(), G21)


def test_zero_weight_sum_and_missing_in_edges():
    df = pd.DataFrame({"from": [1, 2, 3, 1], "to": [2, 3, 2, 4], "weight": [1.0, 0.0, -1.0, 2.0]})
    _, G = orc.directed_graph(df, True)
    _, nodes, op = graph.build_directed(df, True, "from", "to", "weight")
    np.testing.assert_array_equal(op.to_dense(), G)      # 1/0 -> inf -> 0 (SimRank.py:49)
    assert op.dead.sum() >= 2


def test_missing_labels_keep_pandas_count_semantics():
    """`groupby(to)[from].count()` skips missing `from` values (SimRank.py:47); the fast in-degree
    count is only taken when no label is missing."""
    df = pd.DataFrame({"from": [1.0, np.nan, 3.0, 1.0, 2.0], "to": [2.0, 2.0, 2.0, 4.0, 3.0]})
    _, nodes, op = graph.build_directed(df, False, "from", "to", "weight")
    g = dict(zip(nodes, op.g))
    assert g[2.0] == 0.5 and g[4.0] == 1.0 and g[3.0] == 1.0 and g[1.0] == 0.0   # node 2: three rows, two counted
    big = synth.directed_frame(3000, 40000, 0.7, 11)                    # no missing labels: bincount path
    _, G = orc.directed_graph(big, False)
    _, _, op = graph.build_directed(big, False, "from", "to", "weight")
    np.testing.assert_array_equal(op.to_dense(), G)


def test_duplicate_pairs_raise_pivot_error():
    df = pd.DataFrame({"from": [1, 1], "to": [2, 2]})
    with pytest.raises(ValueError, match="Index contains duplicate entries, cannot reshape"):
        graph.build_directed(df, False, "from", "to", "weight")
    with pytest.raises(ValueError, match="Index contains duplicate entries, cannot reshape"):
        graph.build_bipartite(df.rename(columns={"from": "user", "to": "item"}), False, "user", "item", "weight")


def test_missing_column_raises_keyerror():
    with pytest.raises(KeyError):
        graph.build_directed(pd.DataFrame({"a": [1], "b": [2]}), False, "from", "to", "weight")


def test_import_surface_and_constructors():
    from SimRank import SimRank as M
    for name in ["SimRank", "BipartiteSimRank", "SimRankPP", "BipartiteSimRankPP", "AprioriSimRank",
                 "BipartitleAprioriSimRank", "BipartitleSimRank", "BipartitleSimRankPP", "BAR_LENGTH",
                 "update_progress"]:
        assert hasattr(M, name), name
    assert M.BAR_LENGTH == 30
    s = M.SimRank()
    assert s.Nodes == set() and s.Graph.empty
    p = M.SimRankPP()
    assert p.Evidence.empty and p.Weight.empty
    b = M.BipartiteSimRank()
    assert b.NodesGroup1 == set() and b.Graph_N1_N2.empty and b.Graph_N2_N1.empty
    bp = M.BipartiteSimRankPP()
    assert bp.Evidence_N1.empty and bp.Weight_N2.empty and isinstance(bp, M.SimRankPP)
    assert issubclass(M.AprioriSimRank, M.SimRankPP) and issubclass(M.BipartitleAprioriSimRank, M.BipartiteSimRankPP)


def _bar(progress):
    from SimRank.Helper import update_progress
    buf = io.StringIO()
    with redirect_stdout(buf):
        update_progress(progress)
    return buf.getvalue()


def test_progress_bar_text():
    # expected strings follow the format of the reference's Helper.py:16-17
    assert _bar(0.0) == "\rPercent: [" + "-" * 30 + "] 0.0% "
    assert _bar(0.5) == "\rPercent: [" + "#" * 15 + "-" * 15 + "] 50.0% "
    assert _bar(3 / 100) == "\rPercent: [" + "#" * 1 + "-" * 29 + "] 3.0% "
    assert _bar(1) == "\rPercent: [" + "#" * 30 + "] 100% Done...\r\n"
    assert _bar(2.5) == "\rPercent: [" + "#" * 30 + "] 100% Done...\r\n"
    with pytest.raises(ValueError, match="Progress must be float"):
        _bar("0.1")
    with pytest.raises(ValueError, match="Progress below 0"):
        _bar(-0.1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_fit_fails_loudly_without_gpu():
    from SimRank import SimRank as M
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.SimRank().fit(notebook_directed_df(), from_node_column="ORIGIN_AIRPORT_ID",
                        to_node_column="DEST_AIRPORT_ID", verbose=False)


def test_run_loop_semantics():
    from simrank_b200.engine import run_loop
    seq = iter([0.5, 0.2, 0.00005, 0.0])
    applied, conv, last = run_loop(lambda: next(seq), 100, 1e-4, False)
    assert (applied, conv) == (3, True)                 # check happens BEFORE each update
    applied, conv, _ = run_loop(lambda: 0.5, 4, 1e-4, False)
    assert (applied, conv) == (4, False)                # last pair never checked (SimRank.py:129)
    applied, conv, _ = run_loop(lambda: 1 / 0, 5, 1e-4, True, None, (0.0, 0.0))       # empty graph: no update at all
    assert (applied, conv) == (0, True)
    applied, conv, _ = run_loop(lambda: 0.5, 0, 1e-4, False)
    assert (applied, conv) == (0, False)
    applied, conv, _ = run_loop(lambda: 0.5, 5, 1.0, False)
    assert (applied, conv) == (0, True)                 # eps >= 1: |I - 0| <= eps
    pairs = iter([(0.5, 0.00001), (0.00001, 0.5), (0.00001, 0.00001)])
    applied, conv, _ = run_loop(lambda: next(pairs), 100, 1e-4, True)
    assert (applied, conv) == (3, True)                 # both groups must converge (SimRank.py:289)


@pytest.mark.parametrize("all_rows", [False, True])
def test_list_split_plan_partitions_the_neighbour_lists(all_rows):
    """engine.ListSplit (host logic of SRK_CSR_ACCUM): the pieces tile exactly the lists they replace, stay
    inside one range of X rows, respect the piece length, come range by range with the long ones first; a
    piece that is a whole list is marked 'store' (negative slot), everything else 'add'."""
    import torch
    from simrank_b200 import engine
    rng = np.random.default_rng(0)
    M, K, min_deg, piece, ranges = 60, 1000, 100, 64, 3
    deg = rng.integers(0, 40, size=M)
    deg[[3, 10, 49, 0, 5]] = [700, 333, 1000, 100, 0]
    lists = [np.sort(rng.choice(K, size=d, replace=False)) for d in deg]
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = np.concatenate(lists).astype(np.int32)
    sp = engine.ListSplit(torch.from_numpy(indptr), torch.from_numpy(indices), K, min_deg, piece, ranges, all_rows)
    split = (deg >= 1) if all_rows else (deg >= min_deg)
    hubs = np.nonzero(split)[0]
    assert sp.rows == (M if all_rows else hubs.size)
    lo, hi, sl = sp.piece_lo.numpy(), sp.piece_hi.numpy(), sp.piece_slot.numpy()
    span = -(-K // ranges)
    cover, pieces_of, prev = np.zeros(indices.size, int), np.zeros(M, int), (-1, 0)
    for a, b, s in zip(lo, hi, sl):
        slot = s if s >= 0 else -s - 1
        row = slot if all_rows else hubs[slot]
        assert 0 < b - a <= piece and indptr[row] <= a and b <= indptr[row + 1]
        rg = indices[a:b] // span
        assert rg.min() == rg.max()
        assert (rg[0], -(b - a)) >= prev                       # range by range, long pieces first
        prev = (rg[0], -(b - a))
        cover[a:b] += 1
        pieces_of[row] += 1
        assert (s < 0) == ((a, b) == (indptr[row], indptr[row + 1]))
    for r in range(M):
        assert (cover[indptr[r]:indptr[r + 1]] == int(split[r])).all()
    # what is left for the FIRST / FINAL launch: nothing of the split rows, everything of the others
    np.testing.assert_array_equal(sp.row_hi.numpy() - sp.row_lo.numpy(), np.where(split, 0, deg))
    want_slot = np.where(split, np.arange(M) if all_rows else np.cumsum(split) - 1, -1)
    np.testing.assert_array_equal(sp.slot.numpy(), want_slot)
    if all_rows:                                               # rows that are added to, or never written, start from zero
        assert set(sp.zero_slots.tolist()) == set(np.nonzero(pieces_of != 1)[0].tolist())
    else:
        assert sp.zero_slots is None


def test_list_split_plan_thresholds(monkeypatch):
    import torch
    from simrank_b200 import engine
    indptr = torch.tensor([0, 3, 3, 300], dtype=torch.int64)
    indices = torch.arange(300, dtype=torch.int32) % 50
    for k in ("SRK_SPLIT_MIN", "SRK_SPLIT_PIECE", "SRK_SPLIT_RANGE_MB"):
        monkeypatch.delenv(k, raising=False)
    assert engine.ListSplit.plan(indptr, indices, 50, 255) is None            # no row long enough
    sp = engine.ListSplit.plan(indptr, indices, 50, 297)
    assert (sp.min_deg, sp.piece, sp.ranges, sp.rows) == (256, 512, 1, 1)     # the panel fits in L2
    sp = engine.ListSplit.plan(indptr, indices, 138493, 2000)                 # it does not: 142 MB of 1 KB segments
    assert (sp.min_deg, sp.piece, sp.ranges) == (1024, 256, 5) and sp.rows == 0
    monkeypatch.setenv("SRK_SPLIT_MIN", "0")
    assert engine.ListSplit.plan(indptr, indices, 50, 297) is None
    assert engine.ListSplit.plan(indptr, indices, 50, 297, all_rows=True).rows == 3
