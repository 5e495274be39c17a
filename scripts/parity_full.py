#!/usr/bin/env python
"""Full-size parity of the tensor-core (int8 planes) path at BASELINE cfg4 (n = 32768):
K iterations with the fixed-point path vs the float64 CSR path (itself checked against the CPU
oracle to 1e-12 at the sizes the oracle finishes), plus size-independent properties."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from simrank_b200 import engine, graph, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = engine.require_cuda()
frm, to = synth.directed_edges(n, n * 64, 0.5, 4)
op = graph.operator_from_edges(to, frm, n, n)
dop = engine.DeviceOperator(op, dev)
out = {"n": n, "K": K}
ref = engine.DirectedSolver(dop, 0.8, mode="csr")
t = time.time()
d_ref = [ref.step() for _ in range(K)]
out["csr_seconds"] = time.time() - t
for ns in (None, 2, 3):
    sol = engine.DirectedSolver(dop, 0.8, mode="i8", ns=ns)
    ns = ns or "auto"
    t = time.time()
    d = [sol.step() for _ in range(K)]
    torch.cuda.synchronize()
    out[f"i8x{ns}_seconds"] = time.time() - t
    diff = (sol.S - ref.S).abs()
    out[f"i8x{ns}_maxabs_vs_csr_f64"] = float(diff.max())
    out[f"i8x{ns}_maxdiff_seq"] = d
    out[f"i8x{ns}_slices_used"] = sol.half.slices_used
    S = sol.S
    out[f"i8x{ns}_diag_all_one"] = bool((torch.diagonal(S) == 1).all())
    out[f"i8x{ns}_asym"] = float((S - S.T).abs().max())
    out[f"i8x{ns}_range"] = [float(S.min()), float((S - torch.eye(n, device=dev, dtype=S.dtype)).max())]
    del sol, diff
    torch.cuda.empty_cache()
out["csr_maxdiff_seq"] = d_ref
out["csr_asym"] = float((ref.S - ref.S.T).abs().max())
print(json.dumps(out))
