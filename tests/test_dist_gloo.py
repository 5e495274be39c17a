"""World-size 2 and 3 tests of the row-sharded solver's HOST logic on CPU (gloo backend).

The two CUDA launch hooks of simrank_b200.dist.ShardedHalf are replaced by the numpy emulator
of the C ABI (tests/abi_emulator.py); everything else -- shard plans, send/receive block
layout, the K-blocked operand description, bound propagation, the all-to-all and the MAX
all-reduce -- is the product code, and the gathered result must match the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import abi_emulator
from oracle import simrank_oracle as orc
from simrank_b200 import dist as sdist
from simrank_b200 import graph, synth


class CpuHalf(sdist.ShardedHalf):
    """ShardedHalf with host tensors and the ABI emulator instead of the CUDA library."""

    def _init_identity(self):
        for i in range(self.rows):
            self.S[i, self.row0 + i] = 1.0

    def _dense_pattern(self, cols):
        a8 = torch.zeros((self.n_out, self.lda), dtype=torch.uint8)
        rows = np.repeat(np.arange(self.n_out), self.op.deg)
        a8[torch.from_numpy(rows), torch.from_numpy(cols.astype(np.int64))] = 1
        return a8

    def _launch(self, args, name):
        abi_emulator.srk_i8_half(args)

    def _timed(self, name, fn):
        return fn()


class CpuDirected(sdist.ShardedDirectedSolver):
    half_cls = CpuHalf


class CpuBipartite(sdist.ShardedBipartiteSolver):
    half_cls = CpuHalf


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        if case == "directed":
            frm, to = synth.directed_edges(300, 3000, 0.8, 11)
            op = graph.operator_from_edges(to, frm, 300, 300)
            sol = CpuDirected(op, 0.8, ns=3, device=dev)
            diffs = [sol.step() for _ in range(4)]
            S = sol.S.numpy()
            So, _, _ = orc.simrank(op.to_dense(), 0.8, 4, 0.0)
            err = float(np.abs(S - So).max())
            # the all-reduced scalar is the GLOBAL max|dS| on every rank
            ref_diffs = []
            Sa, Sb = np.zeros((300, 300)), np.eye(300)
            for _ in range(4):
                Sa, Sb = Sb, 0.8 * op.to_dense() @ Sb @ op.to_dense().T
                np.fill_diagonal(Sb, 1)
                ref_diffs.append(float(np.abs(Sb - Sa).max()))
            out.put((rank, err, float(np.abs(np.array(diffs) - np.array(ref_diffs)).max()), sol.half.rows))
        else:
            u, i = synth.bipartite_edges(130, 77, 1500, 1.0, 5)
            g1 = np.random.default_rng(1).random(130) * 0.05 + 0.01      # weighted-style row scales
            g2 = np.random.default_rng(2).random(77) * 0.05 + 0.01
            op12 = graph.operator_from_edges(u, i, 130, 77, g1)
            op21 = graph.operator_from_edges(i, u, 77, 130, g2)
            A12 = (op12.to_dense() > 0).astype(np.int64)
            cnt1 = np.minimum(A12 @ A12.T, 255).astype(np.uint8)
            plan1 = sdist.ShardPlan(130, world)
            ev_local = torch.from_numpy(np.ascontiguousarray(cnt1[plan1.start(rank):plan1.stop(rank)]))
            sol = CpuBipartite(op12, op21, 0.8, 0.7, evidence1=ev_local if ev_local.numel() else None, ns=3, device=dev)
            for _ in range(3):
                sol.step()
            S1, S2 = sol.S1.numpy(), sol.S2.numpy()
            E1 = 1 - 0.5 ** (A12 @ A12.T)
            W1, W2 = op12.to_dense(), op21.to_dense()
            s1, s2 = np.eye(130), np.eye(77)
            for _ in range(3):
                s1 = E1 * (0.8 * (W1 @ s2 @ W1.T))
                np.fill_diagonal(s1, 1)
                s2 = 0.7 * (W2 @ s1 @ W2.T)
                np.fill_diagonal(s2, 1)
            out.put((rank, float(max(np.abs(S1 - s1).max(), np.abs(S2 - s2).max())), 0.0, sol.h1.rows))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["directed", "bipartite"])
def test_sharded_solver_matches_oracle(world, case):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    results = sorted(out.get() for _ in range(world))
    assert [r[0] for r in results] == list(range(world))
    for rank, err, derr, rows in results:
        assert err <= 1e-6, (rank, err)                 # fixed-point planes: north-star bound
        assert derr <= 1e-6
    n = 300 if case == "directed" else 130
    assert sum(r[3] for r in results) == n              # the row blocks tile the matrix


def test_shard_plan_and_padded_layout():
    p = sdist.ShardPlan(700, 3)
    assert (p.per, p.blk, p.padded) == (234, 256, 768)
    assert [p.count(r) for r in range(3)] == [234, 234, 232]
    k = np.arange(700)
    pk = p.pad_index(k)
    assert pk[0] == 0 and pk[233] == 233 and pk[234] == 256 and pk[699] == 2 * 256 + 231
    assert len(set(pk.tolist())) == 700
    q = sdist.ShardPlan(32768, 8)
    assert (q.per, q.blk, q.padded) == (4096, 4096, 32768)
    assert np.array_equal(q.pad_index(np.arange(32768)), np.arange(32768))   # no re-layout at cfg4
    e = sdist.ShardPlan(5, 8)
    assert [e.count(r) for r in range(8)] == [1, 1, 1, 1, 1, 0, 0, 0]
