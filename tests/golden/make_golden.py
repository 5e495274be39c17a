#!/usr/bin/env python
"""Generate the committed golden fixtures.  Runs ONLY in the build container
(it reads /root/reference, which does not exist on the GPU box).

Two families of fixtures are written next to this script:

notebook_outputs.json
    Every similarity table printed by the reference's example notebook
    (examples/basic_examples.ipynb, cells 6, 8, 13, 19/20, 22/23, 26/27, 29/30) with the
    "Converged at iteration k" line of the cell that produced it, plus the input edge
    lists reconstructed from those outputs (the notebook's CSV inputs are not shipped;
    derivations in SURVEY.md appendix B).

ref_*.npz
    Outputs of the reference's OWN ``fit`` loops (SimRank/SimRank.py, unmodified source,
    imported from /root/reference) on small seeded graphs.  Two shims are needed on this
    image (pandas 3): ``DataFrame(index=<set>)`` is converted to a list, and -- for the
    directed classes only -- ``_create_graph`` is replaced by a restatement because the
    chained assignment at SimRank.py:52 is a Copy-on-Write no-op under pandas 3 (it
    would leave the graph all-zero).  The bipartite classes run with the first shim
    alone, so their graph build, preprocessing and loops are 100 % reference code.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import re
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


# --------------------------------------------------------------------------- notebook parsing
def _parse_table(text):
    """Parse a (possibly column-wrapped) DataFrame repr into (row_labels, col_labels, values)."""
    blocks, cur = [], []
    for line in text.splitlines():
        if not line.strip():
            if cur:
                blocks.append(cur)
                cur = []
        else:
            cur.append(line.rstrip().rstrip("\\").rstrip())
    if cur:
        blocks.append(cur)
    rows, cols, cols_data = None, [], []
    for blk in blocks:
        header = blk[0].split()
        body = [ln.split() for ln in blk[1:]]
        labels = [b[0] for b in body]
        if rows is None:
            rows = labels
        assert labels == rows
        cols += header
        cols_data.append(np.array([[float(x) for x in b[1:]] for b in body]))
    vals = np.concatenate(cols_data, axis=1)
    return [int(r) for r in rows], [int(c) for c in cols], vals


def notebook_fixtures():
    nb = json.load(open(os.path.join(REF, "examples", "basic_examples.ipynb")))
    cells = nb["cells"]

    def table(i):
        for o in cells[i]["outputs"]:
            if "data" in o and "text/plain" in o["data"]:
                return _parse_table("".join(o["data"]["text/plain"]))
        raise KeyError(i)

    def converged_at(i):
        txt = "".join("".join(o.get("text", "")) for o in cells[i]["outputs"] if o["output_type"] == "stream")
        return int(re.search(r"Converged at iteration (\d+)", txt).group(1))

    out = {}

    def put(name, cls, weighted, fit_cell, tables):
        ent = {"class": cls, "weighted": weighted, "converged_at": converged_at(fit_cell), "tables": []}
        for t in tables:
            r, c, v = table(t)
            ent["tables"].append({"rows": r, "cols": c, "values": v.tolist()})
        out[name] = ent

    put("B1_directed_unweighted", "SimRank", False, 5, [6])
    put("B2_directed_weighted", "SimRank", True, 7, [8])
    put("B3_directed_pp_weighted", "SimRankPP", True, 12, [13])
    put("B4_bipartite_unweighted", "BipartiteSimRank", False, 18, [19, 20])
    put("B5_bipartite_weighted", "BipartiteSimRank", True, 21, [22, 23])
    put("B6_bipartite_pp_unweighted", "BipartiteSimRankPP", False, 25, [26, 27])
    put("B7_bipartite_pp_weighted", "BipartiteSimRankPP", True, 28, [29, 30])

    # ---- reconstructed inputs ------------------------------------
    airports = out["B1_directed_unweighted"]["tables"][0]["rows"]
    a = 12953
    cut = {12889, 14107, 12892}
    edges = [(u, v) for u in airports for v in airports
             if u != v and not ((u == a and v in cut) or (v == a and u in cut))]
    assert len(edges) == 84                                     # "[84 rows x 3 columns]", notebook cell 3
    insum = {12953: 3543, 12889: 3909, 14107: 3482, 11292: 4204, 10397: 4559, 12892: 4500,
             11298: 4295, 12266: 3258, 13930: 4927, 11057: 2862}
    weights = []
    seen = {}
    for (u, v) in edges:
        deg = sum(1 for e in edges if e[1] == v)
        base = insum[v] // deg
        extra = insum[v] - base * deg if v not in seen else 0
        seen[v] = True
        weights.append(base + extra)
    out["inputs_directed"] = {"from": [e[0] for e in edges], "to": [e[1] for e in edges],
                              "flights": weights}

    users = sorted(out["B4_bipartite_unweighted"]["tables"][0]["rows"])
    movies = sorted(out["B4_bipartite_unweighted"]["tables"][1]["rows"])
    rs = np.array([47, 40.5, 42.5, 37.5, 43, 39, 40.5, 39.5, 50, 40])
    cs = np.array([38.5, 42, 46, 44.5, 37, 38.5, 44.5, 42, 41.5, 45])
    assert rs.sum() == cs.sum() == 419.5
    R = np.outer(rs, cs) / rs.sum()                              # any matrix with these marginals
    out["inputs_bipartite"] = {
        "userId": [u for u in users for _ in movies],
        "movieId": [m for _ in users for m in movies],
        "rating": R.reshape(-1).tolist(),
        "note": "row/column sums are positional in label-sorted (pivot) order; compare .values positionally",
    }
    return out


# --------------------------------------------------------------------------- reference via shim
@contextlib.contextmanager
def reference_module():
    """Import the unmodified reference with DataFrame(index=<set>) accepted (pandas<2 behaviour)."""
    orig = pd.DataFrame.__init__

    def init(self, data=None, index=None, columns=None, *a, **k):
        if isinstance(index, (set, frozenset)):
            index = list(index)
        if isinstance(columns, (set, frozenset)):
            columns = list(columns)
        orig(self, data, index, columns, *a, **k)

    pd.DataFrame.__init__ = init
    sys.path.insert(0, REF)
    try:
        from SimRank import SimRank as ref          # noqa: WPS433  (reference package)
        yield ref
    finally:
        pd.DataFrame.__init__ = orig
        sys.path.remove(REF)
        for m in [m for m in sys.modules if m == "SimRank" or m.startswith("SimRank.")]:
            del sys.modules[m]


def _directed_graph_build(self, data, weighted, from_node_column, to_node_column, weight_column):
    """Stand-in for SimRank._create_graph (SimRank.py:24-52) with the same semantics; only the
    CoW-broken row scatter of line 52 is expressed differently (``.loc[rows, cols] = block``)."""
    self.Nodes = set(data[from_node_column].unique()) | set(data[to_node_column].unique())
    order = list(self.Nodes)
    self.Graph = pd.DataFrame(np.zeros((len(order), len(order))), index=order, columns=order)
    if weighted:
        inn = data.groupby(to_node_column)[weight_column].sum().to_frame().rename(columns={weight_column: "inNeighbors"})
    else:
        inn = data.groupby(to_node_column)[from_node_column].count().to_frame().rename(columns={from_node_column: "inNeighbors"})
    data = data.join(inn, on=to_node_column)
    data["_nw"] = (1.0 / data["inNeighbors"]).replace([np.inf, -np.inf], np.nan).fillna(0)
    piv = data.pivot(index=to_node_column, columns=from_node_column, values="_nw").fillna(0)
    self.Graph.loc[piv.index, piv.columns] = piv.values


def _capture(fn):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = fn()
    m = re.search(r"Converged at iteration (\d+)", buf.getvalue())
    return res, (int(m.group(1)) if m else -1), buf.getvalue()


def _rand_directed(rng, n, m, weighted_scale=10.0):
    pairs = set()
    while len(pairs) < m:
        u, v = rng.integers(0, n, 2)
        if u != v:
            pairs.add((int(u), int(v)))
    pairs = sorted(pairs)
    rng.shuffle(pairs)
    df = pd.DataFrame({"from": [100 + 7 * p[0] for p in pairs], "to": [100 + 7 * p[1] for p in pairs]})
    df["weight"] = np.round(rng.lognormal(0.0, 1.0, len(df)) * weighted_scale, 3) + 0.5
    return df


def _rand_bipartite(rng, n1, n2, m):
    pairs = set()
    while len(pairs) < m:
        pairs.add((int(rng.integers(0, n1)), int(rng.integers(0, n2))))
    for a in range(n1):                        # every node appears at least once
        pairs.add((a, int(rng.integers(0, n2))))
    for b in range(n2):
        pairs.add((int(rng.integers(0, n1)), b))
    pairs = sorted(pairs)
    rng.shuffle(pairs)
    df = pd.DataFrame({"user": [1000 + 3 * p[0] for p in pairs], "item": [50 + 11 * p[1] for p in pairs]})
    df["weight"] = rng.choice([0.5, 1, 1.5, 2, 2.5, 3, 3.5, 4, 4.5, 5], len(df))
    return df


def reference_fixtures():
    rng = np.random.default_rng(20240607)
    with reference_module() as ref:
        class DSimRank(ref.SimRank):
            _create_graph = _directed_graph_build

        class DSimRankPP(ref.SimRankPP):
            _create_graph = _directed_graph_build

        class DApriori(ref.AprioriSimRank):
            _create_graph = _directed_graph_build

        cases = []
        # ---- directed: reference loop + preprocessing, restated graph build -------------
        for name, cls, n, m, kw in [
            ("dir_sr_unw", DSimRank, 24, 90, dict(weighted=False, iterations=100, eps=1e-4)),
            ("dir_sr_w", DSimRank, 24, 90, dict(weighted=True, iterations=100, eps=1e-4)),
            ("dir_sr_fixedK", DSimRank, 40, 200, dict(weighted=False, iterations=7, eps=0.0, C=0.6)),
            ("dir_pp_unw", DSimRankPP, 30, 150, dict(weighted=False, iterations=100, eps=1e-4)),
            ("dir_pp_w", DSimRankPP, 30, 150, dict(weighted=True, iterations=12, eps=0.0)),
        ]:
            df = _rand_directed(rng, n, m)
            obj = cls()
            S, k, _ = _capture(lambda: obj.fit(df, verbose=True, **kw))
            cases.append((name, "directed", cls.__mro__[1].__name__, df, kw,
                          {"S": S.values, "labels": np.array(list(S.index))}, k))
        df = _rand_directed(rng, 20, 80)
        obj = DApriori()
        prior = rng.random((len(set(df["from"]) | set(df["to"])),) * 2)
        prior = (prior + prior.T) / 2
        kw = dict(weighted=True, iterations=9, eps=0.0, lbd=0.3)
        S, k, _ = _capture(lambda: obj.fit(df, prior, verbose=True, **kw))
        cases.append(("dir_apriori_w", "directed", "AprioriSimRank", df, kw,
                      {"S": S.values, "labels": np.array(list(S.index)), "prior": prior}, k))

        # ---- bipartite: 100 % reference code ---------------------------------------------
        for name, cls, n1, n2, m, kw in [
            ("bip_sr_unw", ref.BipartiteSimRank, 12, 17, 70, dict(weighted=False, iterations=100, eps=1e-4)),
            ("bip_sr_w", ref.BipartiteSimRank, 12, 17, 70, dict(weighted=True, iterations=100, eps=1e-4, C1=0.7, C2=0.9)),
            ("bip_sr_fixedK", ref.BipartiteSimRank, 21, 9, 80, dict(weighted=False, iterations=6, eps=0.0)),
            ("bip_pp_sq_unw", ref.BipartiteSimRankPP, 14, 14, 60, dict(weighted=False, iterations=100, eps=1e-4)),
            ("bip_pp_sq_w", ref.BipartiteSimRankPP, 14, 14, 60, dict(weighted=True, iterations=8, eps=0.0)),
        ]:
            df = _rand_bipartite(rng, n1, n2, m)
            obj = cls()
            (S1, S2), k, _ = _capture(lambda: obj.fit(df, verbose=True, **kw))
            cases.append((name, "bipartite", cls.__name__, df, kw,
                          {"S1": S1.values, "S2": S2.values,
                           "sorted1": np.array(sorted(df["user"].unique())),
                           "sorted2": np.array(sorted(df["item"].unique())),
                           "ref_labels1": np.array(list(S1.index)), "ref_labels2": np.array(list(S2.index))}, k))
        # ---- bipartite SimRank++ with priors (SimRank.py:457-493): 100 % reference code.  Priors are
        # given positionally in the label-sorted (pivot) order of the matrices they are added to.
        for name, n, m, sym, kw in [
            ("bip_apriori_sq_w", 14, 60, True, dict(weighted=True, iterations=8, eps=0.0, lbd1=0.3, lbd2=0.6)),
            ("bip_apriori_sq_unw", 11, 45, False, dict(weighted=False, iterations=100, eps=1e-4, C1=0.7, C2=0.9)),
        ]:
            df = _rand_bipartite(rng, n, n, m)
            p1, p2 = rng.random((n, n)), rng.random((n, n))
            if sym:
                p1, p2 = (p1 + p1.T) / 2, (p2 + p2.T) / 2
            obj = ref.BipartitleAprioriSimRank()
            (S1, S2), k, _ = _capture(lambda: obj.fit(df, p1, p2, verbose=True, **kw))
            cases.append((name, "bipartite", "BipartitleAprioriSimRank", df, kw,
                          {"S1": S1.values, "S2": S2.values, "prior1": p1, "prior2": p2,
                           "sorted1": np.array(sorted(df["user"].unique())),
                           "sorted2": np.array(sorted(df["item"].unique())),
                           "ref_labels1": np.array(list(S1.index)), "ref_labels2": np.array(list(S2.index))}, k))
        # n1 != n2 raises for the Apriori variant too (Evidence_N1 in the group-2 update, SimRank.py:491)
        df = _rand_bipartite(rng, 9, 13, 40)
        try:
            _capture(lambda: ref.BipartitleAprioriSimRank().fit(df, rng.random((9, 9)), rng.random((13, 13)),
                                                                verbose=False))
            raised_apriori = ""
        except ValueError as e:
            raised_apriori = str(e)
        assert "broadcast" in raised_apriori

        # n1 != n2: the reference raises (SimRank.py:423) -- recorded as a fact
        df = _rand_bipartite(rng, 9, 13, 40)
        try:
            _capture(lambda: ref.BipartiteSimRankPP().fit(df, verbose=False))
            raised = ""
        except ValueError as e:
            raised = str(e)
        assert "broadcast" in raised

    index = {}
    for name, family, cls, df, kw, arrays, k in cases:
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"),
                            **{c: df[c].to_numpy() for c in df.columns}, **arrays)
        index[name] = {"family": family, "class": cls, "kwargs": kw, "converged_at": k,
                       "columns": list(df.columns)}
    index["_bipartite_pp_rectangular_raises"] = raised
    index["_bipartite_apriori_rectangular_raises"] = raised_apriori
    return index


if __name__ == "__main__":
    nb = notebook_fixtures()
    json.dump(nb, open(os.path.join(HERE, "notebook_outputs.json"), "w"), indent=1)
    idx = reference_fixtures()
    json.dump(idx, open(os.path.join(HERE, "ref_index.json"), "w"), indent=1)
    print("wrote", len(nb), "notebook entries and", len([k for k in idx if not k.startswith("_")]), "reference cases")
