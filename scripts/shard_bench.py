"""Strong scaling of ONE solver sharded over the visible GPUs inside the solve call (single process, chb_solver_shard).

    python scripts/shard_bench.py <workload> [k] [steps]

Prints one JSON line per device count (1, 2, 4, ... up to the visible GPUs): ms per .solve(b, x) of the workload's k
columns (CUDA events on device 0's legacy stream, the stream the call is joined to; L2 flushed between steps), the
speed-up over one GPU and the bandwidth model T(g) = (F + R / g) / BW of SURVEY.md 7.3-5 beside it.
"""
import ctypes
import json
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
    _, d_dtype, d_k = problems.resolve_workload(name)
    k = int(sys.argv[2]) if len(sys.argv) > 2 else d_k
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    dtype = np.float32 if d_dtype == "f" else np.float64
    p = problems.build_problem(name)
    perm, cp, ri, v = bench.get_factor(name, p)
    lib = _capi.load()
    cls = CholeskySolverF if dtype == np.float32 else CholeskySolverD
    h = ctypes.c_void_p()
    _capi.check(lib.chb_init(0))
    _capi.check(lib.chb_solver_create_from_factor(ctypes.byref(h), _capi.F32 if dtype == np.float32 else _capi.F64, len(perm), len(ri),
                                                  cp.ctypes.data, ri.ctypes.data, v.ctypes.data, perm.ctypes.data, 0))
    s = cls._from_handle(h, len(perm), 0)
    n, z, es = len(perm), len(ri), np.dtype(dtype).itemsize
    torch.cuda.set_device(0)
    b = torch.rand((n, k), dtype=torch.float32 if dtype == np.float32 else torch.float64, device="cuda:0")
    x = torch.zeros_like(b)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    x_one, t_one = None, None
    counts = [g for g in (1, 2, 4, 8) if g <= torch.cuda.device_count()]
    for g in counts:
        t = time.time()
        s.shard(list(range(1, g)))
        t_shard = time.time() - t
        info = s.shard_info()
        for _ in range(3):
            s.solve(b, x)
        torch.cuda.synchronize()
        ts = []
        for i in range(steps):
            flush.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.solve(b, x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        s.check_error()
        xs = x.cpu().numpy()
        if x_one is None:
            x_one, t_one = xs, float(np.median(ts))
        ms = float(np.median(ts))
        factor_bytes = 2 * z * (4 + es) + 2 * 4 * (n + 1) + 4 * 4 * n
        rhs_bytes = 4 * n * k * es
        model = (factor_bytes + rhs_bytes) / (factor_bytes + rhs_bytes / g)
        print(json.dumps({"workload": bench.workload_description(name, p, k, dtype), "devices": info["devices"], "ms_per_solve": ms,
                          "columns_per_s": k / (ms * 1e-3), "speedup": t_one / ms, "bandwidth_model_speedup": model,
                          "replication": info["replication"], "replicated_MB": info["bytes"] / 1e6,
                          "replication_s": info["seconds"], "shard_call_s": t_shard,
                          "bit_identical_to_one_gpu": bool(np.array_equal(xs, x_one))}), flush=True)


if __name__ == "__main__":
    main()
