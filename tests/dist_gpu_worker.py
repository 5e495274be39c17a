"""torchrun worker for tests/test_gpu_dist.py: runs the drop-in classes under NCCL with one rank
per GPU and checks the row-sharded result against the oracle on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import simrank_oracle as orc  # noqa: E402
from simrank_b200 import synth  # noqa: E402
from SimRank import SimRank as M  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    # directed, uneven shards (n = 1000 is not a multiple of world * 128)
    df = synth.directed_frame(1000, 20000, 0.8, 21)
    nodes, So, ko, co = orc.fit_directed(df, iterations=5, eps=0.0)
    S = M.SimRank(mode="i8").fit(df, iterations=5, eps=0.0, verbose=False)
    assert list(S.index) == nodes
    worst = max(worst, float(np.abs(S.to_numpy() - So).max()))
    # convergence decision is global: same iteration count as the oracle
    nodes, So, ko, co = orc.fit_directed(df, iterations=50, eps=1e-4)
    obj = M.SimRank(mode="i8")
    S = obj.fit(df, iterations=50, eps=1e-4, verbose=False)
    assert (obj.fit_info_.applied, obj.fit_info_.converged) == (ko, co), (obj.fit_info_, ko, co)
    worst = max(worst, float(np.abs(S.to_numpy() - So).max()))
    # gather="local": every rank keeps its own row block (all columns)
    Sl = M.SimRank(mode="i8", gather="local").fit(df, iterations=50, eps=1e-4, verbose=False)
    pos = {lab: i for i, lab in enumerate(nodes)}
    mine = [pos[lab] for lab in Sl.index]
    assert list(Sl.columns) == nodes and mine == list(range(mine[0], mine[0] + len(mine))) if mine else True
    if mine:
        worst = max(worst, float(np.abs(Sl.to_numpy() - So[mine]).max()))
    tot = torch.tensor([len(mine)], device="cuda")
    dist.all_reduce(tot)
    assert int(tot.item()) == len(nodes)
    # SimRank++ (evidence rows are sharded with S)
    df = synth.directed_frame(1500, 30000, 1.0, 22, weights="lognormal")
    nodes, So, _, _ = orc.fit_directed(df, kind="simrank_pp", weighted=True, iterations=4, eps=0.0)
    S = M.SimRankPP(mode="i8").fit(df, weighted=True, iterations=4, eps=0.0, verbose=False)
    worst = max(worst, float(np.abs(S.to_numpy() - So).max()))
    # bipartite, rectangular
    df = synth.bipartite_frame(700, 300, 20000, 1.0, 23)
    l1, l2, S1o, S2o, _, _ = orc.fit_bipartite(df, weighted=True, iterations=4, eps=0.0)
    S1, S2 = M.BipartiteSimRank(mode="i8").fit(df, weighted=True, iterations=4, eps=0.0, verbose=False)
    worst = max(worst, float(np.abs(S1.to_numpy() - S1o).max()), float(np.abs(S2.to_numpy() - S2o).max()))
    # BASELINE cfg5 at 1/16 scale: BipartiteSimRankPP, n1 != n2 (Evidence_N2 for group 2), weighted
    df = synth.config_frame("cfg5", scale=1 / 16)
    l1, l2, S1o, S2o, _, _ = orc.fit_bipartite(df, kind="simrank_pp", weighted=True, iterations=3, eps=0.0)
    S1, S2 = M.BipartitleSimRankPP(mode="i8").fit(df, weighted=True, iterations=3, eps=0.0, verbose=False)
    assert list(S1.index) == l1 and list(S2.index) == l2
    worst = max(worst, float(np.abs(S1.to_numpy() - S1o).max()), float(np.abs(S2.to_numpy() - S2o).max()))
    t = torch.tensor([worst], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DIST_OK world={world} max_abs_err={t.item():.3e}")
    assert t.item() <= 1e-6, t.item()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
