"""World-size 2, 3 and 4 tests of the row-sharded solver's HOST logic on CPU (gloo backend).

The CUDA launch hooks of simrank_b200.dist.ShardedHalf are replaced by the numpy emulator of the
C ABI (tests/abi_emulator.py); everything else -- shard plans, the assignment of the blocks of the
symmetric update to ranks, staging layouts, pointer offsets, bound propagation, the all-to-alls
and the MAX all-reduce -- is the product code (the StagedExchange strategy; on the GPU box the
same offsets address peer memory), and the gathered result must match the oracle."""
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

    def _dense_pattern(self):
        a8 = torch.zeros((self.n_out, self.lda), dtype=torch.uint8)
        rows = np.repeat(np.arange(self.n_out), self.op.deg)
        a8[torch.from_numpy(rows), torch.from_numpy(self.op.indices.astype(np.int64))] = 1
        return a8

    def _launch(self, args, name):
        abi_emulator.srk_x2_half(args)

    def _launch_slice(self, ns):
        abi_emulator.srk_slice_rows_max_f64(self.S.data_ptr(), self.ld, self.rows, self.n_out, self.row0, ns,
                                            self.planes.data_ptr(), self.ldp, self.planes.stride(0),
                                            self.bound_vec.data_ptr())

    def _timed(self, name, fn):
        return fn()


class CpuCsrHalf(sdist.ShardedCsrHalf):
    """ShardedCsrHalf with host tensors and the emulated CSR half-product."""

    def _init_identity(self):
        for i in range(self.rows):
            self.S[i, self.row0 + i] = 1.0

    def _launch_csr(self, row_begin, row_end, x_ptr, ldx, L, out_ptr, ldo, epi):
        assert ldx >= L and ldo >= row_end - row_begin            # the library's own argument check
        abi_emulator.srk_csr_half_f64(self.op.indptr, self.op.indices, np.asarray(self.op.g, dtype=np.float64),
                                      self.n_out, row_begin, row_end, x_ptr, ldx, L, out_ptr, ldo, epi, self.n_in)

    def _timed(self, name, fn):
        return fn()


class CpuCsr16Half(sdist.ShardedCsr16Half):
    """ShardedCsr16Half (uint16 gather: quantiser, split neighbour lists, ACCUM + FINISH / FINISH_FIRST, the float64
    fallback update) with host tensors and the emulated library."""

    _init_identity = CpuCsrHalf._init_identity
    _launch_csr = CpuCsrHalf._launch_csr
    _dense_pattern = CpuHalf._dense_pattern
    _launch = CpuHalf._launch

    def _library(self):
        return abi_emulator.Library

    def _timed(self, name, fn):
        return fn()


class CpuDirected(sdist.ShardedDirectedSolver):
    half_cls = CpuHalf
    csr_half_cls = CpuCsrHalf
    csr16_half_cls = CpuCsr16Half


class CpuBipartite(sdist.ShardedBipartiteSolver):
    half_cls = CpuHalf
    csr_half_cls = CpuCsrHalf
    csr16_half_cls = CpuCsr16Half


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
            sol = CpuDirected(op, 0.8, ns=None, device=dev)
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
        elif case == "drivers":
            # the routing layer the drop-in classes call (simrank_b200.drivers) with the device pieces
            # replaced: operator -> device is a stub, the solvers use the emulated kernels
            import types
            from simrank_b200 import drivers
            drivers._device_op = lambda op, device=None: types.SimpleNamespace(device=dev)
            sdist.ShardedDirectedSolver.half_cls, sdist.ShardedDirectedSolver.csr_half_cls = CpuHalf, CpuCsrHalf
            sdist.ShardedBipartiteSolver.half_cls, sdist.ShardedBipartiteSolver.csr_half_cls = CpuHalf, CpuCsrHalf
            frm, to = synth.directed_edges(260, 2600, 0.8, 13)
            op = graph.operator_from_edges(to, frm, 260, 260)
            G = op.to_dense()
            worst = 0.0
            for mode, want_mode in ((None, "i8"), ("i8", "i8"), ("csr", "csr")):
                sol = drivers.directed_solver(op, 0.8, mode=mode, device=dev)
                assert sol.mode == want_mode
                applied, conv, last = drivers.run_loop(sol.step, 3, 0.0, False)
                res = drivers.collect(sol, [list(range(260))], "local")
                (a, b), = res.rows
                So, _, _ = orc.simrank(G, 0.8, 3, 0.0)
                worst = max(worst, float(np.abs(res.mats[0].numpy() - So[a:b]).max()))
                full = drivers.collect(sol, [list(range(260))], "all")
                worst = max(worst, float(np.abs(full.mats[0].numpy() - So).max()))
            # SimRank++ with the evidence of the operator's own pattern, bipartite, through the same layer
            u, i = synth.bipartite_edges(120, 70, 1400, 1.0, 6)
            op12, op21 = graph.operator_from_edges(u, i, 120, 70), graph.operator_from_edges(i, u, 70, 120)
            ev1, ev2 = drivers.EvidenceMatrix(op12, dev), drivers.EvidenceMatrix(op21, dev)
            sol = drivers.bipartite_solver(op12, op21, 0.8, 0.8, evidence1=ev1, evidence2=ev2, device=dev)
            assert sol.mode == "i8" and sol.h1.evidence_from_pattern and sol.h2.evidence_from_pattern
            drivers.run_loop(sol.step, 3, 0.0, True)
            W1, W2 = op12.to_dense(), op21.to_dense()
            s1o, s2o, _, _ = orc.bipartite_simrank_pp(W1, W2, orc.evidence(W1), orc.evidence(W2), 0.8, 0.8, 3, 0.0)
            res = drivers.collect(sol, [list(range(120)), list(range(70))], "all")
            worst = max(worst, float(np.abs(res.mats[0].numpy() - s1o).max()), float(np.abs(res.mats[1].numpy() - s2o).max()))
            with pytest.raises(NotImplementedError, match="symmetric prior"):
                drivers.directed_solver(op, 0.8, prior=np.triu(np.ones((260, 260))), lbd=0.5, device=dev)
            out.put((rank, worst, 0.0, sol.h1.rows))
        elif case.startswith("directed_csr16"):
            # the fixed-point CSR path, what 'auto' picks for sparse graphs on N GPUs: hub rows so that the split
            # neighbour lists are exercised; "_fused" keeps the fused launches instead of ACCUM + streaming pass
            if case.endswith("_fused"):
                os.environ.update(SRK_FINAL_VIA_ACCUM="0", SRK_FIRST_VIA_ACCUM="0")
            os.environ.update(SRK_SPLIT_MIN="40", SRK_SPLIT_PIECE="16")
            rng = np.random.default_rng(14)
            mask = rng.random((230, 230)) < 0.03
            mask[7] = rng.random(230) < 0.6                           # two hub rows
            mask[101] = rng.random(230) < 0.4
            mask[55] = False                                          # and an empty one
            op = graph.operator_from_edges(*np.nonzero(mask), 230, 230)
            sol = CpuDirected(op, 0.8, mode="csr16", device=dev)
            assert sol.mode == "csr16" and sol.half.split is not None
            assert (sol.half.split_all is not None) == (not case.endswith("_fused"))
            diffs = [sol.step() for _ in range(5)]
            G = op.to_dense()
            So, _, _ = orc.simrank(G, 0.8, 5, 0.0)
            err = float(np.abs(sol.S.numpy() - So).max())
            assert 2 in sol.half.slices_used                           # the uint16 updates really ran
            Sa, Sb, ref_diffs = np.zeros((230, 230)), np.eye(230), []
            for _ in range(5):
                Sa, Sb = Sb, 0.8 * G @ Sb @ G.T
                np.fill_diagonal(Sb, 1)
                ref_diffs.append(float(np.abs(Sb - Sa).max()))
            out.put((rank, err, float(np.abs(np.array(diffs) - np.array(ref_diffs)).max()), sol.half.rows))
        elif case == "bipartite_csr16":
            # BipartiteSimRankPP shape (n1 != n2), evidence taken from the pattern counts, fixed-point CSR path
            os.environ.update(SRK_SPLIT_MIN="30", SRK_SPLIT_PIECE="16", SRK_SPLIT_RANGE_MB="0.05")
            u, i = synth.bipartite_edges(130, 77, 1500, 1.0, 5)
            op12, op21 = graph.operator_from_edges(u, i, 130, 77), graph.operator_from_edges(i, u, 77, 130)
            sol = CpuBipartite(op12, op21, 0.8, 0.7, mode="csr16", device=dev, evidence1_from_pattern=True,
                               evidence2_from_pattern=True)
            assert sol.mode == "csr16" and sol.h2.split is not None and sol.h2.split.ranges > 1
            for _ in range(3):
                sol.step()
            W1, W2 = op12.to_dense(), op21.to_dense()
            s1o, s2o, _, _ = orc.bipartite_simrank_pp(W1, W2, orc.evidence(W1), orc.evidence(W2), 0.8, 0.7, 3, 0.0)
            err = float(max(np.abs(sol.S1.numpy() - s1o).max(), np.abs(sol.S2.numpy() - s2o).max()))
            out.put((rank, err, 0.0, sol.h1.rows))
        elif case == "directed_csr":
            # negative weight sums: rows of G with a negative scale (SimRank.py:45,49) -- float64 CSR path
            frm, to = synth.directed_edges(210, 1800, 0.8, 12)
            gsc = np.random.default_rng(3).random(210) * 0.2 - 0.05
            op = graph.operator_from_edges(to, frm, 210, 210, gsc)
            prior = np.random.default_rng(4).random((210, 210))
            prior = 0.5 * (prior + prior.T)          # sharded fits need a symmetric prior (drivers._require_symmetric_prior)
            plan = sdist.ShardPlan(210, world)
            sol = CpuDirected(op, 0.8, prior=torch.from_numpy(prior[plan.start(rank):plan.stop(rank)].copy()), lbd=0.25,
                              mode=None, device=dev)
            assert sol.mode == "csr"
            diffs = [sol.step() for _ in range(4)]
            G = op.to_dense()
            Sa, Sb, ref_diffs = np.zeros((210, 210)), np.eye(210), []
            for _ in range(4):
                Sa, Sb = Sb, 0.75 * (0.8 * G @ Sb @ G.T) + 0.25 * prior
                np.fill_diagonal(Sb, 1)
                ref_diffs.append(float(np.abs(Sb - Sa).max()))
            err = float(np.abs(sol.S.numpy() - Sb).max())
            out.put((rank, err, float(np.abs(np.array(diffs) - np.array(ref_diffs)).max()), sol.half.rows))
        elif case == "bipartite_csr":
            u, i = synth.bipartite_edges(130, 77, 1500, 1.0, 5)
            g1 = np.random.default_rng(1).random(130) * 0.05 - 0.01     # some negative row scales
            g2 = np.random.default_rng(2).random(77) * 0.05 + 0.01
            op12 = graph.operator_from_edges(u, i, 130, 77, g1)
            op21 = graph.operator_from_edges(i, u, 77, 130, g2)
            A12 = (op12.to_dense() != 0).astype(np.int64)
            cnt = np.minimum(A12 @ A12.T, 255).astype(np.uint8)
            plan1 = sdist.ShardPlan(130, world)
            ev = torch.zeros((plan1.count(rank), 144), dtype=torch.uint8)
            ev[:, :130] = torch.from_numpy(cnt[plan1.start(rank):plan1.stop(rank)].copy())
            sol = CpuBipartite(op12, op21, 0.8, 0.7, evidence1=ev, mode="auto", device=dev)
            assert sol.mode == "csr"
            for _ in range(3):
                sol.step()
            S1, S2 = sol.S1.numpy(), sol.S2.numpy()
            E1 = 1 - 0.5 ** cnt.astype(np.float64)
            W1, W2 = op12.to_dense(), op21.to_dense()
            s1, s2 = np.eye(130), np.eye(77)
            for _ in range(3):
                s1 = E1 * (0.8 * (W1 @ s2 @ W1.T))
                np.fill_diagonal(s1, 1)
                s2 = 0.7 * (W2 @ s1 @ W2.T)
                np.fill_diagonal(s2, 1)
            out.put((rank, float(max(np.abs(S1 - s1).max(), np.abs(S2 - s2).max())), 0.0, sol.h1.rows))
        else:
            u, i = synth.bipartite_edges(130, 77, 1500, 1.0, 5)
            g1 = np.random.default_rng(1).random(130) * 0.05 + 0.01      # weighted-style row scales
            g2 = np.random.default_rng(2).random(77) * 0.05 + 0.01
            op12 = graph.operator_from_edges(u, i, 130, 77, g1)
            op21 = graph.operator_from_edges(i, u, 77, 130, g2)
            A12 = (op12.to_dense() > 0).astype(np.int64)
            sol = CpuBipartite(op12, op21, 0.8, 0.7, ns=3, device=dev, evidence1_from_pattern=True)
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


@pytest.mark.parametrize("world,case", [(2, "directed"), (3, "directed"), (4, "directed"), (2, "bipartite"),
                                        (3, "bipartite"), (4, "bipartite"), (2, "directed_csr"),
                                        (3, "bipartite_csr"), (4, "bipartite_csr"), (2, "drivers"),
                                        (2, "directed_csr16"), (3, "directed_csr16"), (2, "directed_csr16_fused"),
                                        (3, "bipartite_csr16")])
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
    n = {"directed": 300, "directed_csr": 210, "drivers": 120, "directed_csr16": 230, "directed_csr16_fused": 230}.get(case, 130)
    assert sum(r[3] for r in results) == n              # the row blocks tile the matrix


def test_sharded_solver_refuses_what_the_fixed_point_path_cannot_hold():
    op = graph.operator_from_edges([0, 1, 2], [1, 2, 0], 3, 3, g=np.array([0.5, -0.25, 1.0]))
    with pytest.raises(ValueError, match="non-negative"):
        sdist._sharded_mode("i8", op)
    assert sdist._sharded_mode(None, op) == "csr"                  # auto: the float64 path takes it
    ok = graph.operator_from_edges([0, 1], [1, 0], 2, 2)
    assert sdist._sharded_mode(None, ok) == "i8" and sdist._sharded_mode("csr", ok) == "csr"
    with pytest.raises(ValueError, match="unknown mode"):
        sdist._sharded_mode("tf32", ok)


def test_sharded_fit_needs_a_symmetric_prior():
    from simrank_b200 import drivers
    p = np.random.default_rng(0).random((5, 5))
    with pytest.raises(NotImplementedError, match="symmetric prior"):
        drivers._require_symmetric_prior(p)
    drivers._require_symmetric_prior(None, p + p.T)


def test_shard_plan_and_block_assignment():
    p = sdist.ShardPlan(700, 3)
    assert p.per == 240 and [p.count(r) for r in range(3)] == [240, 240, 220]
    q = sdist.ShardPlan(32768, 8)
    assert q.per == 4096 and q.start(7) == 28672
    e = sdist.ShardPlan(5, 8)
    assert [e.count(r) for r in range(8)] == [5, 0, 0, 0, 0, 0, 0, 0]
    # every unordered pair of row blocks is computed exactly once, by one of its two owners
    for n, world in ((700, 3), (32768, 8), (1000, 4), (300, 2), (5, 8), (4096, 5)):
        plan = sdist.ShardPlan(n, world)
        cover = np.zeros((n, n), dtype=np.int32)
        for rank in range(world):
            for (pp, j_lo, j_hi, r_lo, r_hi, mirror) in sdist.pair_tasks(plan, rank, True):
                rows = slice(plan.start(rank) + r_lo, plan.start(rank) + r_hi)
                cols = slice(plan.start(pp) + j_lo, plan.start(pp) + j_hi)
                cover[rows, cols] += 1
                if mirror:
                    cover[cols, rows] += 1
                else:
                    assert pp == rank
        assert (cover == 1).all(), (n, world)
        # balanced: no rank computes more than its share plus one split block
        work = [sum((j_hi - j_lo) * (r_hi - r_lo) * (0.5 if pp == rank else 1.0)
                    for (pp, j_lo, j_hi, r_lo, r_hi, _) in sdist.pair_tasks(plan, rank, True)) for rank in range(world)]
        if n >= 256 * world:
            assert max(work) <= 1.35 * (sum(work) / world), (n, world, work)
