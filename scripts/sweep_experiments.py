"""Timing experiments for the sweep kernels on one GPU (run under gpurun).

    python scripts/sweep_experiments.py <workload> <dtype f|d> "<k list>" "<CHB_SWEEP_FLAGS list>" "<ctas/SM list>"
"""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
import problems  # noqa: E402
from cholespy_b200 import CholeskySolverD, CholeskySolverF, _capi  # noqa: E402


def main():
    name = sys.argv[1]
    dt = sys.argv[2]
    ks = [int(v) for v in sys.argv[3].split(",")]
    flags = sys.argv[4].split(",") if len(sys.argv) > 4 else ["0"]
    ctas = sys.argv[5].split(",") if len(sys.argv) > 5 else [""]
    builds = sys.argv[6].split(",") if len(sys.argv) > 6 else [""]
    p = problems.build_problem(name)
    perm, cp, ri, v = bench.get_factor(name, p)
    lib = _capi.load()
    A = problems.to_scipy_csc(p)
    touched = set()
    for build in builds:
        for key in touched:  # a build must not inherit the previous build's settings
            os.environ.pop(key, None)
        touched = {kv.split("=")[0] for kv in filter(None, build.split("+"))}
        for kv in filter(None, build.split("+")):
            key, val = kv.split("=")
            os.environ[key] = val
        print("== build env:", build or "(defaults)", flush=True)
        run_one(lib, p, A, perm, cp, ri, v, dt, ks, flags, ctas)


def run_one(lib, p, A, perm, cp, ri, v, dt, ks, flags, ctas):
    cls = CholeskySolverF if dt == "f" else CholeskySolverD
    h = ctypes.c_void_p()
    t = time.time()
    _capi.check(lib.chb_init(0))
    _capi.check(lib.chb_solver_create_from_factor(ctypes.byref(h), _capi.F32 if dt == "f" else _capi.F64, len(perm), len(ri),
                                                  cp.ctypes.data, ri.ctypes.data, v.ctypes.data, perm.ctypes.data, 0))
    s = cls._from_handle(h, len(perm), 0)
    print("image built in %.1f s" % (time.time() - t), s.info(), flush=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for k in ks:
        b = torch.rand((p["n"], k), dtype=torch.float32 if dt == "f" else torch.float64, device="cuda")
        x = torch.zeros_like(b)
        for fl in flags:
            for c in ctas:
                os.environ["CHB_SWEEP_FLAGS"] = fl
                if c:
                    os.environ["CHB_SWEEP_CTAS_PER_SM"] = c
                else:
                    os.environ.pop("CHB_SWEEP_CTAS_PER_SM", None)
                for _ in range(2):
                    s.solve(b, x)
                torch.cuda.synchronize()
                ts = []
                for i in range(5):
                    flush.fill_(i)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    s.solve(b, x)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                s.check_error()
                xs = x[:, :2].double().cpu().numpy()
                bs = b[:, :2].double().cpu().numpy()
                res = float((np.linalg.norm(A @ xs - bs, axis=0) / np.linalg.norm(bs, axis=0)).max())
                print("k=%d flags=%s ctas/SM=%s grid=%d: median %.3f ms  min %.3f ms  (%.0f cols/s)  residual %.2e" % (
                    k, fl, c or "max", s.info()["grid"], float(np.median(ts)), min(ts), k / (np.median(ts) * 1e-3), res), flush=True)


if __name__ == "__main__":
    main()
