#!/usr/bin/env python
"""BASELINE cfg1-cfg3 through the drop-in classes on one B200: wall time of the second fit (the first
one warms the allocator), its stages, parity against the CPU oracle and the oracle's own time on
this host.  cfg2 also times the SimRank++ preprocessing that replaces `_cal_Evidence` /
`_cal_Weight` (SimRank.py:311-337).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import simrank_oracle as orc  # noqa: E402
from simrank_b200 import synth  # noqa: E402
from SimRank import SimRank as M  # noqa: E402


MODES = ("csr", "csr16", "i8", "auto")


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t


def main():
    out = {}
    for name, cls, okind, bip, weighted in (("cfg1", M.SimRank, "simrank", False, True),
                                            ("cfg2", M.SimRankPP, "simrank_pp", False, True),
                                            ("cfg3", M.BipartiteSimRank, "simrank", True, True)):
        cfg = synth.CONFIGS[name]
        K = cfg["iterations"]
        df = synth.config_frame(name)
        row = {"class": cfg["cls"], "iterations": K}
        for mode in MODES:
            obj = cls(mode=None if mode == "auto" else mode)
            kw = dict(weighted=weighted, iterations=K, eps=0.0, verbose=False)
            obj.fit(df, **kw)                                     # warm-up
            res, dt = timed(lambda: obj.fit(df, **kw))
            row[mode] = {"seconds": round(dt, 4), "stages_s": {k: round(v, 4) for k, v in obj.fit_timings_.items()},
                         "ms_per_iteration": round(1e3 * obj.fit_timings_["iterate"] / K, 3),
                         "mode_used": obj.fit_info_.mode}
            row[mode]["result"] = res
        t = time.perf_counter()
        if bip:
            _, _, S1o, S2o, _, _ = orc.fit_bipartite(df, kind=okind, weighted=weighted, iterations=K, eps=0.0)
            want = (S1o, S2o)
        else:
            _, So, _, _ = orc.fit_directed(df, kind=okind, weighted=weighted, iterations=K, eps=0.0)
            want = (So,)
        row["oracle_cpu_seconds"] = round(time.perf_counter() - t, 3)
        for mode in MODES:
            got = row[mode].pop("result")
            got = got if isinstance(got, tuple) else (got,)
            row[mode]["max_abs_vs_oracle"] = max(float(np.abs(g.to_numpy() - w).max()) for g, w in zip(got, want))
        if name == "cfg2":                                        # preprocessing kernels on their own
            obj = M.SimRankPP(mode="i8")
            obj._create_graph(df, weighted, "from", "to", "weight")
            for _ in range(2):
                W, t_w = timed(lambda: obj._cal_Weight(obj._graph_op, False))
                E, t_e = timed(lambda: (lambda e: (e.counts, e)[1])(obj._cal_Evidence(obj._graph_op, False)))
            row["weight_seconds"], row["evidence_counts_seconds"] = round(t_w, 5), round(t_e, 5)
            G = obj._graph_op.to_dense()
            t = time.perf_counter()
            Eo = orc.evidence(G)
            row["oracle_evidence_cpu_seconds"] = round(time.perf_counter() - t, 3)
            row["evidence_bit_exact"] = bool(np.array_equal(np.asarray(E), Eo))
            t = time.perf_counter()
            Wo = orc.weight(G)
            row["oracle_weight_cpu_seconds"] = round(time.perf_counter() - t, 3)
            row["weight_max_abs"] = float(np.abs(np.asarray(W) - Wo).max())
        out[name] = row
    print(json.dumps(out))


if __name__ == "__main__":
    main()
