"""Per-dependency-level summary of a trace written by scripts/trace_front.py.

    python scripts/trace_levels.py gpurun_out/trace_front_<...>.npz [fwd|bwd]
"""
import sys

import numpy as np

g = np.load(sys.argv[1])
keys = sys.argv[2:] or ["fwd", "bwd"]
for key in keys:
    tr = g[key].astype(np.int64)
    lvl, by, loc = g[key + "_level"], g[key + "_bytes"], g[key + "_local"]
    t0 = tr[:, 0].min()
    start, end, wait = tr[:, 0] - t0, tr[:, 2] - t0, tr[:, 4]
    print(key, "tasks", len(lvl))
    print(" lvl     n  local  firstStart   lastEnd   meanDur   maxDur  meanWait (us)  meanBytes  sumMB")
    for l in range(int(lvl.max()) + 1):
        m = lvl == l
        print(" %3d %5d %6d %10.1f %9.1f %9.1f %8.1f %9.1f %10.0f %7.1f" % (
            l, m.sum(), loc[m].sum(), start[m].min() / 1e3, end[m].max() / 1e3, (end - start)[m].mean() / 1e3,
            (end - start)[m].max() / 1e3, wait[m].mean() / 1e3, by[m].mean(), by[m].sum() / 1e6))
