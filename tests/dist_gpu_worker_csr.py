"""torchrun worker for tests/test_gpu_zz_dist_csr.py: the row-sharded float64 CSR solver
(simrank_b200.dist.ShardedCsrHalf) under NCCL, one rank per GPU, against the oracle."""
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
    # explicit mode="csr": same graphs as the tensor-core worker, float64 tolerance
    df = synth.directed_frame(1000, 20000, 0.8, 21)
    nodes, So, ko, co = orc.fit_directed(df, iterations=50, eps=1e-4)
    obj = M.SimRank(mode="csr")
    S = obj.fit(df, iterations=50, eps=1e-4, verbose=False)
    assert obj.fit_info_.mode == "csr" and list(S.index) == nodes
    assert (obj.fit_info_.applied, obj.fit_info_.converged) == (ko, co), (obj.fit_info_, ko, co)
    worst = max(worst, float(np.abs(S.to_numpy() - So).max()))
    # negative weight sums: 'auto' must take the float64 path (the planes hold non-negative values)
    dfw = synth.directed_frame(700, 9000, 1.0, 31, weights="lognormal")
    dfw["weight"] = dfw["weight"] - 0.5                       # two nodes end up with a negative weight sum
    nodes, So, _, _ = orc.fit_directed(dfw, weighted=True, iterations=4, eps=0.0)
    obj = M.SimRank()
    S = obj.fit(dfw, weighted=True, iterations=4, eps=0.0, verbose=False)
    assert obj.fit_info_.mode == "csr"
    scale = max(1.0, float(np.abs(So).max()))
    worst = max(worst, float(np.abs(S.to_numpy() - So).max()) / scale)
    # SimRank++ (uint8 evidence counts, sharded with S) and the rectangular bipartite alternation
    df = synth.directed_frame(1500, 30000, 1.0, 22, weights="lognormal")
    nodes, So, _, _ = orc.fit_directed(df, kind="simrank_pp", weighted=True, iterations=4, eps=0.0)
    S = M.SimRankPP(mode="csr").fit(df, weighted=True, iterations=4, eps=0.0, verbose=False)
    worst = max(worst, float(np.abs(S.to_numpy() - So).max()))
    df = synth.config_frame("cfg5", scale=1 / 32)
    l1, l2, S1o, S2o, _, _ = orc.fit_bipartite(df, kind="simrank_pp", weighted=True, iterations=3, eps=0.0)
    S1, S2 = M.BipartitleSimRankPP(mode="csr").fit(df, weighted=True, iterations=3, eps=0.0, verbose=False)
    assert list(S1.index) == l1 and list(S2.index) == l2
    worst = max(worst, float(np.abs(S1.to_numpy() - S1o).max()), float(np.abs(S2.to_numpy() - S2o).max()))
    # the same solver with the gathers in uint16 fixed point (mode="csr16"): 1e-6 bound
    worst16 = 0.0
    df = synth.directed_frame(1000, 20000, 0.8, 21)
    nodes, So, ko, co = orc.fit_directed(df, iterations=50, eps=1e-4)
    obj = M.SimRank(mode="csr16")
    S = obj.fit(df, iterations=50, eps=1e-4, verbose=False)
    assert obj.fit_info_.mode == "csr16" and (obj.fit_info_.applied, obj.fit_info_.converged) == (ko, co)
    worst16 = max(worst16, float(np.abs(S.to_numpy() - So).max()))
    df = synth.config_frame("cfg5", scale=1 / 32)
    l1, l2, S1o, S2o, _, _ = orc.fit_bipartite(df, kind="simrank_pp", weighted=True, iterations=3, eps=0.0)
    S1, S2 = M.BipartitleSimRankPP(mode="csr16").fit(df, weighted=True, iterations=3, eps=0.0, verbose=False)
    worst16 = max(worst16, float(np.abs(S1.to_numpy() - S1o).max()), float(np.abs(S2.to_numpy() - S2o).max()))
    t = torch.tensor([worst, worst16], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DIST_CSR_OK world={world} max_abs_err={t[0].item():.3e} csr16_max_abs_err={t[1].item():.3e}")
    assert t[0].item() <= 1e-12, t[0].item()
    assert t[1].item() <= 1e-6, t[1].item()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
