"""Dumps the per-task timeline of one solve (debug; run under gpurun with a built library).

    python scripts/trace_sweep.py <workload> <f|d> <k> [out.npz]
"""
import ctypes
import os
import sys

import numpy as np

os.environ["CHB_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import problems  # noqa: E402
from cholespy_b200 import CholeskySolverD, CholeskySolverF, _capi  # noqa: E402


def main():
    name, dt, k = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out", "trace_%s_%s_k%d.npz" % (name, dt, k))
    p = problems.build_problem(name)
    perm, cp, ri, v = bench.get_factor(name, p)
    lib = _capi.load()
    cls = CholeskySolverF if dt == "f" else CholeskySolverD
    h = ctypes.c_void_p()
    _capi.check(lib.chb_init(0))
    _capi.check(lib.chb_solver_create_from_factor(ctypes.byref(h), _capi.F32 if dt == "f" else _capi.F64, len(perm), len(ri),
                                                  cp.ctypes.data, ri.ctypes.data, v.ctypes.data, perm.ctypes.data, 0))
    s = cls._from_handle(h, len(perm), 0)
    info = s.info()
    b = torch.rand((p["n"], k), dtype=torch.float32 if dt == "f" else torch.float64, device="cuda")
    x = torch.zeros_like(b)
    for _ in range(3):
        s.solve(b, x)
    torch.cuda.synchronize()
    res = {}
    for upper, key in ((0, "fwd"), (1, "bwd")):
        nt = info[key + "_tasks"]
        buf = np.zeros(8 * nt, np.uint64)
        got = lib.chb_solver_get_trace(h, upper, buf.ctypes.data, buf.size)
        assert got == buf.size, (got, buf.size)
        res[key] = buf.reshape(nt, 8)
        t0 = res[key][:, 0].min()
        print(key, "tasks", nt, "span %.3f ms" % ((res[key][:, 2].max() - t0) / 1e6), flush=True)
    # the task table of the same (deterministic) schedule, built on the host: bytes, level, first record type
    hs = ctypes.c_void_p()
    _capi.check(lib.chb_schedule_create(ctypes.byref(hs), _capi.F32 if dt == "f" else _capi.F64, len(perm), len(ri), cp.ctypes.data,
                                        ri.ctypes.data, v.ctypes.data, perm.ctypes.data, -1))
    for upper, key in ((0, "fwd"), (1, "bwd")):
        slab, tasks, order = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        sinfo = (ctypes.c_int64 * 10)()
        lib.chb_schedule_sweep(hs, upper, ctypes.byref(slab), ctypes.byref(tasks), ctypes.byref(order), sinfo, 10)
        nt = int(sinfo[0])
        td = np.ctypeslib.as_array(ctypes.cast(tasks, ctypes.POINTER(ctypes.c_uint32)), shape=(nt, 4)).copy()
        sl = np.ctypeslib.as_array(ctypes.cast(slab, ctypes.POINTER(ctypes.c_uint32)), shape=(int(sinfo[1]) // 4,))
        off = (td[:, 0].astype(np.int64) | (td[:, 1].astype(np.int64) << 32)) // 4
        res[key + "_bytes"] = td[:, 2]
        res[key + "_level"] = td[:, 3]
        res[key + "_first"] = sl[off] & 0xff
    lib.chb_schedule_destroy(hs)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, **res)
    print("wrote", out)


if __name__ == "__main__":
    main()
