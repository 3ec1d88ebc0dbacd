"""Per-level timeline of a panel-sweep trace written by scripts/trace_sweep.py.

    python scripts/trace_sweep_levels.py gpurun_out/trace_<workload>_<f|d>_k<k>.npz
"""
import sys

import numpy as np

g = np.load(sys.argv[1])
for key in ("fwd", "bwd"):
    tr = g[key].astype(np.int64)
    lvl = g[key + "_level"].astype(np.int64)
    byt = g[key + "_bytes"].astype(np.int64) & 0x0fffffff
    ran = tr[:, 0] > 0   # tasks of the local-task / wide-level kernels leave no trace record
    t0 = tr[ran, 0].min()
    st, en = np.where(ran, (tr[:, 0] - t0) / 1e3, 0.0), np.where(ran, (tr[:, 2] - t0) / 1e3, 0.0)   # us
    span = en.max()
    print("%s: %d tasks, %d levels, span %.3f ms, sum of task durations / (span x CTAs): see utilisation" % (key, len(st), lvl.max() + 1, span / 1e3))
    print("  level  tasks   MB    first start   last end (us)   level span   mean task us   p99 task us")
    for L in range(lvl.max() + 1):
        m = (lvl == L) & ran
        if not m.any():
            continue
        d = en[m] - st[m]
        print("  %5d %6d %6.1f %12.1f %12.1f %12.1f %12.1f %12.1f" % (L, m.sum(), byt[m].sum() / 1e6, st[m].min(), en[m].max(),
                                                                  en[m].max() - st[m].min(), d.mean(), np.percentile(d, 99)))
    order = np.argsort(np.where(ran, st, np.inf))[:int(ran.sum())]
    q = [st[order[int(f * (len(order) - 1))]] for f in (0.25, 0.5, 0.75, 0.9, 0.99, 1.0)]
    print("  task start times at 25/50/75/90/99/100%% of the tasks: %s us" % ", ".join("%.0f" % v for v in q))
    bq = np.cumsum(byt[order]) / byt[order].sum()
    for f in (0.5, 0.9, 0.99):
        print("  %.0f%% of the slab bytes started by %.0f us" % (100 * f, st[order[min(np.searchsorted(bq, f), len(order) - 1)]]))
