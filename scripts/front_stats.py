"""Prints the shape of the FRONT sweep schedule of a workload (host only, no GPU).

    python scripts/front_stats.py <workload> [f|d]
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
import problems  # noqa: E402
from cholespy_b200 import _capi  # noqa: E402
from oracle.oracle import FRONT_INFO_KEYS  # noqa: E402


def main():
    name = sys.argv[1]
    dt = sys.argv[2] if len(sys.argv) > 2 else "f"
    p = problems.build_problem(name)
    perm, cp, ri, v = bench.get_factor(name, p)
    lib = _capi.load()
    h = ctypes.c_void_p()
    t = time.time()
    _capi.check(lib.chb_front_schedule_create(ctypes.byref(h), 0 if dt == "f" else 1, len(perm), len(ri), cp.ctypes.data,
                                              ri.ctypes.data, v.ctypes.data, perm.ctypes.data))
    print("front schedule built in %.1f s" % (time.time() - t))
    es = 4 if dt == "f" else 8
    for upper in (0, 1):
        slab, tasks, order = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        info = (ctypes.c_int64 * 27)()
        lib.chb_front_schedule_sweep(h, upper, ctypes.byref(slab), ctypes.byref(tasks), ctypes.byref(order), info, 27)
        raw = list(info)
        info = dict(zip(FRONT_INFO_KEYS, raw))
        print("  local critical path per task: groups (subtrees, upper A, upper B, contributions) = %s, rows = %s" % (
            [round(x / max(info["n_local_tasks"], 1), 1) for x in raw[19:23]], [round(x / max(info["n_local_tasks"], 1), 1) for x in raw[23:27]]))
        nt = info["n_tasks"]
        td = np.ctypeslib.as_array(ctypes.cast(tasks, ctypes.POINTER(ctypes.c_uint32)), shape=(nt, 4))
        nbytes, level = td[:, 2].astype(np.int64), td[:, 3]
        print("backward" if upper else "forward", info)
        print("  slab/nnz bytes ratio %.3f; stored/nnz %.3f; task bytes: median %d p90 %d max %d; W rows x 512B = %.1f MB" % (
            info["slab_bytes"] / (info["n_values"] * es), info["n_stored_values"] / info["n_values"], np.median(nbytes),
            np.percentile(nbytes, 90), nbytes.max(), info["n_w_rows"] * 512 / 1e6))
        per_level = np.bincount(level)
        print("  task levels %d; tasks per level: first 8 %s; levels with 1 task: %d" % (
            len(per_level), per_level[:8].tolist(), int((per_level == 1).sum())))
    lib.chb_front_schedule_destroy(h)


if __name__ == "__main__":
    main()
