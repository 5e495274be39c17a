#!/usr/bin/env python
"""Summarise SRK_X2_TRACE dumps (scripts/gpu_trace.sh): per-tile mainloop / epilogue durations, the
wait of the MMA issuer for a free accumulator and how far the CTA pairs drift apart."""
import glob
import re
import sys

import numpy as np


def report(path):
    m = re.search(r"\.m(\d)\.ns(\d)\.c(\d+)\.t(\d+)\.bin$", path)
    mode, ns, C, T = (int(x) for x in m.groups())
    a = np.fromfile(path, dtype=np.uint64).reshape(C, T, 4).astype(np.int64)
    valid = a[:, :, 0] > 0
    t0 = a[:, :, 0][valid].min()
    us = np.where(a > 0, (a - t0) / 1e3, np.nan)
    ml, ep = us[:, :, 1] - us[:, :, 0], us[:, :, 3] - us[:, :, 2]
    gap = us[:, 1:, 0] - us[:, :-1, 1]
    total = np.nanmax(us[:, :, 3])
    out = [f"{path}: mode {('MID', 'FINAL', 'COUNTS')[mode]} ns={ns} pairs={C} tiles/pair={T} total {total / 1e3:.2f} ms"]
    for lo, hi in ((2, 12), (T // 2 - 5, T // 2 + 5), (T - 14, T - 4)):
        sl = slice(lo, hi)
        spread = np.nanmax(us[:, sl, 0], axis=0) - np.nanmin(us[:, sl, 0], axis=0)
        out.append(f"  tiles {lo:3d}-{hi:3d}: mainloop {np.nanmean(ml[:, sl]):6.1f} us  epilogue {np.nanmean(ep[:, sl]):6.1f} us"
                   f"  issuer wait {np.nanmean(gap[:, lo - 1:hi - 1]):5.1f} us  start spread {np.nanmean(spread):6.0f} us")
    return "\n".join(out)


if __name__ == "__main__":
    for pat in sys.argv[1:]:
        for f in sorted(glob.glob(pat)):
            print(report(f))
