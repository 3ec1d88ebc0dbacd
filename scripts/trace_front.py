"""Per-task timeline of one solve on the front path (debug; run under gpurun with a built library).

    python scripts/trace_front.py <workload> <f|d> <k> [out.npz]

Prints, per sweep: span, CTA utilisation (sum of task durations / (CTAs x span)), share of task time spent waiting
for staged rows, idle gaps between consecutive tasks of a CTA, and the same split by task kind (local / dense)."""
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

FR_LOADX, FR_LROWS = 1, 7


def main():
    name, dt, k = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out", "trace_front_%s_%s_k%d.npz" % (name, dt, k))
    p = problems.build_problem(name)
    perm, cp, ri, v = bench.get_factor(name, p)
    lib = _capi.load()
    cls = CholeskySolverF if dt == "f" else CholeskySolverD
    code = _capi.F32 if dt == "f" else _capi.F64
    h = ctypes.c_void_p()
    _capi.check(lib.chb_init(0))
    _capi.check(lib.chb_solver_create_from_factor(ctypes.byref(h), code, len(perm), len(ri), cp.ctypes.data, ri.ctypes.data,
                                                  v.ctypes.data, perm.ctypes.data, 0))
    s = cls._from_handle(h, len(perm), 0)
    info = s.info()
    b = torch.rand((p["n"], k), dtype=torch.float32 if dt == "f" else torch.float64, device="cuda")
    x = torch.zeros_like(b)
    for _ in range(3):
        s.solve(b, x)
    torch.cuda.synchronize()
    s.check_error()
    grid = s.info()["grid"]
    hs = ctypes.c_void_p()
    _capi.check(lib.chb_front_schedule_create(ctypes.byref(hs), code, len(perm), len(ri), cp.ctypes.data, ri.ctypes.data,
                                              v.ctypes.data, perm.ctypes.data))
    res = {}
    for upper, key in ((0, "fwd"), (1, "bwd")):
        nt = info[key + "_tasks"]
        buf = np.zeros(8 * nt + 16 * 4096, np.uint64)
        got = lib.chb_solver_get_trace(h, upper, buf.ctypes.data, buf.size)
        assert got == buf.size, (got, buf.size)
        tr = buf[:8 * nt].reshape(nt, 8).astype(np.int64)
        prof = buf[8 * nt:].reshape(4096, 16)[:grid, :13].astype(np.float64)
        names = ["queue wait", "page wait", "staged wait", "MMA", "EPI", "ASM", "BAR", "REL", "LROWS", "LSTORE", "end barrier",
                 "flag release", "other"]
        tot = prof.sum()
        print("%s: compute warp 0 time by activity (all CTAs): " % key + ", ".join(
            "%s %.1f%%" % (nm, 100 * prof[:, i].sum() / tot) for i, nm in enumerate(names)) + "; mean total per CTA %.2f ms" % (prof.sum(axis=1).mean() / 1e6))
        slab, tasks, order = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        sinfo = (ctypes.c_int64 * 19)()
        lib.chb_front_schedule_sweep(hs, upper, ctypes.byref(slab), ctypes.byref(tasks), ctypes.byref(order), sinfo, 19)
        assert int(sinfo[0]) == nt
        td = np.ctypeslib.as_array(ctypes.cast(tasks, ctypes.POINTER(ctypes.c_uint32)), shape=(nt, 4)).copy()
        od = np.ctypeslib.as_array(ctypes.cast(order, ctypes.POINTER(ctypes.c_int32)), shape=(nt,)).copy()
        sl = np.ctypeslib.as_array(ctypes.cast(slab, ctypes.POINTER(ctypes.c_uint32)), shape=(int(sinfo[1]) // 4,))
        off = (td[:, 0].astype(np.int64) | (td[:, 1].astype(np.int64) << 32)) // 4
        # a local task's third record is FR_LROWS (after the page directory and its FR_LOADX)
        first_len = (sl[off + 4] >> 8 & 0xffff).astype(np.int64) * 4
        local = (sl[off + 4 + first_len] & 0xff) == FR_LROWS
        ticket = np.empty(nt, np.int64)
        ticket[od] = np.arange(nt)
        t0 = tr[:, 0].min()
        start, end, cta, first, wait = tr[:, 0] - t0, tr[:, 2] - t0, tr[:, 1], tr[:, 3] - t0, tr[:, 4]
        span = end.max()
        dur = end - start
        print("%s: tasks %d (local %d), span %.3f ms, grid %d, utilisation %.1f%%, waiting for staged rows %.1f%% of task time" % (
            key, nt, int(local.sum()), span / 1e6, grid, 100.0 * dur.sum() / (grid * span), 100.0 * wait.sum() / max(dur.sum(), 1)))
        for nm, m in (("local", local), ("dense", ~local)):
            if m.any():
                print("   %s: n %d, duration mean %.1f us p50 %.1f p90 %.1f max %.1f; wait mean %.1f us; start->first rows mean %.1f us; bytes mean %.0f" % (
                    nm, int(m.sum()), dur[m].mean() / 1e3, np.median(dur[m]) / 1e3, np.percentile(dur[m], 90) / 1e3, dur[m].max() / 1e3,
                    wait[m].mean() / 1e3, np.maximum(first[m] - start[m], 0).mean() / 1e3, td[m, 2].mean()))
        ls, lf, ld = tr[:, 5] - t0, tr[:, 6] - t0, tr[:, 7] - t0
        tsel = (int(os.environ.get("CHB_FRONT_DEBUG", "0")) >> 8) & 3
        for nm, m in (("local", local), ("dense", ~local)):
            if m.any() and tsel:
                mm = m & (tr[:, 5] > 0) & (tr[:, 6] > 0) & (tr[:, 7] > 0)
                print("   %s loader events (sel %d): w5 -> w6 %.2f us, w6 -> w7 %.2f us; compute start - w5 %.2f us; compute end - w7 %.2f us" % (
                    nm, tsel, (lf - ls)[mm].mean() / 1e3, (ld - lf)[mm].mean() / 1e3, (start - ls)[mm].mean() / 1e3, (end - ld)[mm].mean() / 1e3))
            elif m.any():
                print("   %s loader: busy per task mean %.1f us (start -> first staged issue %.1f us); loader start -> compute start %.1f us; loader done -> compute end %.1f us" % (
                    nm, (ld - ls)[m].mean() / 1e3, np.where(lf[m] > 0, lf[m] - ls[m], 0).mean() / 1e3, (start - ls)[m].mean() / 1e3,
                    (end - ld)[m].mean() / 1e3))
        # gaps between consecutive tasks of one CTA
        o = np.lexsort((start, cta))
        same = cta[o][1:] == cta[o][:-1]
        gaps = (start[o][1:] - end[o][:-1])[same]
        print("   gaps between tasks of a CTA: mean %.1f us p50 %.1f p90 %.1f max %.1f; total %.1f%% of CTA time" % (
            gaps.mean() / 1e3, np.median(gaps) / 1e3, np.percentile(gaps, 90) / 1e3, gaps.max() / 1e3, 100.0 * gaps.sum() / (grid * span)))
        # progress: when was the ticket at 25/50/75/100 % started
        qs = [int(q * (nt - 1)) for q in (0.25, 0.5, 0.75, 1.0)]
        print("   ticket start times at 25/50/75/100%%: %s ms; level of the task graph %d" % (
            ", ".join("%.2f" % (start[od[q]] / 1e6) for q in qs), int(td[:, 3].max()) + 1))
        res[key] = tr
        res[key + "_bytes"], res[key + "_level"], res[key + "_local"], res[key + "_ticket"] = td[:, 2], td[:, 3], local, ticket
    lib.chb_front_schedule_destroy(hs)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, **res)
    print("wrote", out)


if __name__ == "__main__":
    main()
