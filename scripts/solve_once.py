"""A few solves of one workload on cuda:0 (target for ncu / compute-sanitizer runs under gpurun).

    python scripts/solve_once.py <workload> <f|d> <k> [n_solves]
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import problems  # noqa: E402
from cholespy_b200 import CholeskySolverD, CholeskySolverF, _capi  # noqa: E402


def main():
    name, dt, k = sys.argv[1], sys.argv[2], int(sys.argv[3])
    n_solves = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    p = problems.build_problem(name)
    perm, cp, ri, v = bench.get_factor(name, p)
    lib = _capi.load()
    cls = CholeskySolverF if dt == "f" else CholeskySolverD
    h = ctypes.c_void_p()
    _capi.check(lib.chb_init(0))
    _capi.check(lib.chb_solver_create_from_factor(ctypes.byref(h), _capi.F32 if dt == "f" else _capi.F64, len(perm), len(ri),
                                                  cp.ctypes.data, ri.ctypes.data, v.ctypes.data, perm.ctypes.data, 0))
    s = cls._from_handle(h, len(perm), 0)
    b = torch.rand((p["n"], k), dtype=torch.float32 if dt == "f" else torch.float64, device="cuda")
    x = torch.zeros_like(b)
    for _ in range(n_solves):
        s.solve(b, x)
        torch.cuda.synchronize()
    s.check_error()
    A = problems.to_scipy_csc(p)
    xs = x[:, :2].double().cpu().numpy()
    bs = b[:, :2].double().cpu().numpy()
    print("residual %.3e" % float((np.linalg.norm(A @ xs - bs, axis=0) / np.linalg.norm(bs, axis=0)).max()))


if __name__ == "__main__":
    main()
