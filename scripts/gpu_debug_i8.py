#!/usr/bin/env python
"""Diagnostics for the tcgen05 path: one-hot probes that expose WHICH operand element every
output element picked up (swizzle / descriptor / TMEM-lane mistakes show up as permutations)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from simrank_b200 import _lib, engine

dev = engine.require_cuda()
lib = _lib.load()
print("cc", lib.srk_device_cc(), "i8 supported", lib.srk_i8_supported(), torch.cuda.get_device_name(0))


def counts(P, A):
    R, K = P.shape
    N = A.shape[0]
    ldk = engine._round_up(K, 128)
    ldn = engine._round_up(N, 16)
    Pd = torch.zeros((R, ldk), dtype=torch.uint8, device=dev); Pd[:, :K] = torch.from_numpy(P)
    Ad = torch.zeros((N, ldk), dtype=torch.uint8, device=dev); Ad[:, :K] = torch.from_numpy(A)
    out = torch.full((R, ldn), 77, dtype=torch.uint8, device=dev)
    a = _lib.I8Args()
    a.mode, a.ns, a.R, a.N, a.K = _lib.SRK_I8_COUNTS, 1, R, N, K
    a.in_planes, a.ld_in, a.in_plane_stride = Pd.data_ptr(), ldk, R * ldk
    a.A8, a.lda = Ad.data_ptr(), ldk
    a.out_planes, a.ld_outp, a.out_plane_stride = out.data_ptr(), ldn, R * ldn
    rc = lib.srk_i8_half(C.byref(a), engine._stream())
    if rc:
        print("rc", rc, lib.srk_last_error())
        return None
    torch.cuda.synchronize()
    return out[:, :N].cpu().numpy()


rng = np.random.default_rng(0)
for (R, N, K) in [(128, 256, 128), (128, 256, 32), (128, 256, 256), (128, 256, 1024), (256, 512, 128), (100, 70, 50),
                  (300, 300, 300)]:
    P = rng.integers(0, 200, (R, K), dtype=np.uint8)
    perm = rng.integers(0, K, N)
    A = np.zeros((N, K), dtype=np.uint8)
    A[np.arange(N), perm] = 1
    got = counts(P, A)
    if got is None:
        continue
    want = P[:, perm]
    ok = np.array_equal(got, want)
    print(f"one-hot R={R} N={N} K={K}: {'OK' if ok else 'MISMATCH'}  frac_equal={np.mean(got == want):.4f}")
    if not ok:
        bad = np.argwhere(got != want)
        print("  first mismatches (r, j, got, want, perm[j]):")
        for r, j in bad[:12]:
            # where does the value we got live in row r of P?
            src = np.nonzero(P[r] == got[r, j])[0][:6]
            print(f"   r={r} j={j} got={got[r, j]} want={want[r, j]} k_want={perm[j]} k_candidates={src.tolist()}")
        rows_bad = np.unique(bad[:, 0]); cols_bad = np.unique(bad[:, 1])
        print("  bad rows:", rows_bad[:20].tolist(), "... n=", len(rows_bad), " bad cols:", cols_bad[:20].tolist(), "... n=", len(cols_bad))
    # dense random check
    A2 = (rng.random((N, K)) < 0.3).astype(np.uint8)
    P2 = (rng.random((R, K)) < 0.3).astype(np.uint8)
    got = counts(P2, A2)
    want = np.minimum(P2.astype(np.int64) @ A2.astype(np.int64).T, 255)
    print(f"random  R={R} N={N} K={K}: {'OK' if np.array_equal(got, want) else 'MISMATCH'} frac_equal={np.mean(got == want):.4f}")
