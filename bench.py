#!/usr/bin/env python
"""Benchmark of the cholespy solve path on B200: RHS-columns solved per second (L + L' sweeps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4] [--impl reference|reference-gpu]

One step = one `.solve(b, x)` of the workload's system (forward + backward sweep over all RHS columns).
Default workload: BASELINE.json's config the metric is quoted on -- cfg4, icosphere(9) I+10L,
CholeskySolverF, 128 RHS (RHS columns sharded over the ranks when N > 1: strong scaling, the default).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.

Timing: CUDA events on the legacy default stream (the stream the kernels are launched on) around each
solve, L2 flushed between steps, max over ranks.  `e2e` goes through the public API with pinned host
buffers (H2D + solve + D2H inside the timed region).  `cpu_baseline` times the reference's own CHOLMOD
solve_cpu on ONE core on a bounded column sample (oracle/_ref, the reference sources compiled unmodified; the C
oracle port where that library is absent); `--impl reference` times the same solve_cpu fanned over all host cores;
`--impl reference-gpu` times the reference's own GPU kernels (kernels/kernels.cu:64-202 through its solve_cuda,
src/cholesky_solver.cpp:415-432) on the same B200 -- "the kernel to beat" of BASELINE.md section 3.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import problems  # noqa: E402

METRIC = "rhs_columns_solved_per_sec"
UNIT = "RHS-columns/s"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workload + factor cache (CHOLMOD's simplicial factorisation of cfg4 takes minutes; both arms and all
# N reuse one on-disk copy per box)
# ------------------------------------------------------------------------------------------------
def cache_dir():
    for d in (os.environ.get("CHB_CACHE"), "/tmp/chb_cache", os.path.join(ROOT, ".cache")):
        if not d:
            continue
        try:
            os.makedirs(d, exist_ok=True)
            return d
        except OSError:
            pass
    return None


def workload_description(name, p, k, dtype):
    return "%s: n=%d, %s input, CholeskySolver%s, %d RHS" % (name, p["n"], p["fmt"], "F" if dtype == np.float32 else "D", k)


def get_factor(name, p):
    """(perm, l_colptr, l_rowidx, l_vals) for problem p; cached on disk."""
    from cholespy_b200 import _capi
    from helpers import factorize
    d = cache_dir()
    key = hashlib.sha1(("%s|%d|%d|%.17g" % (name, p["n"], len(p["x"]), float(np.sum(p["x"])))).encode()).hexdigest()[:16]
    path = os.path.join(d, "factor_%s_%s.npz" % (name, key)) if d else None
    if path and os.path.exists(path):
        t = time.time()
        g = np.load(path)
        out = (g["perm"], g["l_colptr"], g["l_rowidx"], g["l_vals"])
        log("factor cache hit %s (%.1f s)" % (path, time.time() - t))
        return out
    t = time.time()
    out = factorize(_capi.load(), p["n"], p["ii"], p["jj"], p["x"], p["fmt"])
    log("CHOLMOD factorisation: %.1f s, nnz(L)=%d" % (time.time() - t, len(out[2])))
    if path:
        try:
            tmp = path + ".tmp%d.npz" % os.getpid()
            np.savez(tmp, perm=out[0], l_colptr=out[1], l_rowidx=out[2], l_vals=out[3])
            os.replace(tmp, path)
        except OSError as e:
            log("factor cache not written:", e)
    return out


def algorithmic_bytes(n, z, k, s):
    """factor both triangles once + row pointers + level order + permutation + RHS traffic."""
    return 2 * z * (4 + s) + 2 * 4 * (n + 1) + 2 * 4 * n + 2 * 4 * n + 4 * n * k * s


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (CHOLMOD solve_cpu)
# ------------------------------------------------------------------------------------------------
_REF_SOLVE = None  # set before the worker pool forks: workers inherit the factorised reference solver


def _ref_worker(b):
    _REF_SOLVE(b)
    return b.shape[1]


def run_reference(args, name, p, k, dtype):
    """The reference's own CPU implementation of the path: CholeskySolver<Float>::solve_cpu (CHOLMOD
    simplicial solve, reference src/cholesky_solver.cpp:434-463) compiled unmodified into oracle/_ref.
    That solve is single-threaded by construction, so "all the host threads it can use" = one forked
    worker per core, each solving its own block of 4 RHS columns (CHOLMOD's internal column blocking)
    of one step's sample concurrently with the same factor."""
    import multiprocessing as mp
    global _REF_SOLVE
    from oracle import oracle
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t = time.time()
    if oracle.reference_available():
        kind = "reference"
        ref = oracle.ReferenceSolver(p["n"], p["ii"], p["jj"], p["x"], p["fmt"], dtype, cpu=True)
        log("reference constructor (CHOLMOD analyze+factorize): %.1f s" % (time.time() - t))
        _REF_SOLVE = lambda b: ref.solve_cpu(b)  # noqa: E731
    else:
        kind = "port"
        perm, cp, ri, v = get_factor(name, p)
        _REF_SOLVE = lambda b: oracle.solve(cp, ri, v, perm, b)  # noqa: E731
    cores = max(1, min(os.cpu_count() or 1, args.ref_procs, (k + 3) // 4))
    sample_k = min(k, 4 * cores)
    rng = np.random.default_rng(45)
    b = rng.random((p["n"], sample_k)).astype(dtype)
    parts = [np.ascontiguousarray(b[:, c:c + 4]) for c in range(0, sample_k, 4)]
    pool = None
    if cores > 1:
        try:
            pool = mp.get_context("fork").Pool(cores)
        except (OSError, ValueError) as e:
            log("fork pool unavailable (%s): single process" % e)
            cores = 1
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if pool:
            pool.map(_ref_worker, parts, chunksize=1)
        else:
            for part in parts:
                _REF_SOLVE(part)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
        if sum(times) > args.ref_budget_s and len(times) >= 1:
            break
    if pool:
        pool.close()
    ms = 1e3 * float(np.mean(times))
    value = sample_k / (ms * 1e-3)
    sample = "%d of %d RHS columns per step (4 per worker), %d timed steps, %d worker processes sharing one factor (CHOLMOD's simplicial solve is single-threaded)" % (
        sample_k, k, len(times), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
        "config": {"workload": workload_description(name, p, k, dtype), "sample_columns": sample_k},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# reference GPU arm: the reference's own kernels on this B200
# ------------------------------------------------------------------------------------------------
def run_reference_gpu(args, name, p, k, dtype):
    """Times CholeskySolver<Float>::solve_cuda of the REFERENCE (oracle/_ref: src/*.cpp compiled unmodified; its
    shipped sm_30 PTX JIT-ed by the driver, or kernels/kernels.cu compiled for sm_100a if the driver refuses the PTX)
    with CUDA events on the legacy stream, L2 flushed between steps.  The reference's analysis kernels race on a B200
    (profiles/r01_reference_analysis_race.txt): a schedule that is not even a permutation is rejected and the
    constructor retried; the caller should run this arm under a timeout (an invalid schedule can spin forever)."""
    import torch
    from oracle import oracle
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not oracle.reference_available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libcholespy_ref.so is not built"}), flush=True)
        return
    torch.cuda.set_device(0)
    t = time.time()
    ref, tries = None, 0
    while ref is None and tries < 4:
        tries += 1
        try:
            cand = oracle.ReferenceSolver(p["n"], p["ii"], p["jj"], p["x"], p["fmt"], dtype, cpu=False)
        except RuntimeError:
            cand = oracle.ReferenceSolver(p["n"], p["ii"], p["jj"], p["x"], p["fmt"], dtype, cpu=False, use_sm100_cubin=True)
        lo, up, _ = cand.schedule()
        n = p["n"]
        if np.array_equal(np.sort(lo), np.arange(n)) and np.array_equal(np.sort(up), np.arange(n)):
            ref = cand
        else:
            log("reference analysis produced an invalid level order (race), retrying")
            cand.close()
    if ref is None:
        print(json.dumps({"impl": "reference-gpu", "unavailable": "reference analysis never produced a valid level order"}), flush=True)
        return
    log("reference GPU constructor (CHOLMOD + its analysis kernels): %.1f s, %d attempt(s)" % (time.time() - t, tries))
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    b = torch.rand((p["n"], k), dtype=tdt, device="cuda", generator=torch.Generator(device="cuda").manual_seed(45))
    x = torch.zeros_like(b)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(max(args.warmup, 1)):
        ref.solve_cuda(k, b.data_ptr(), x.data_ptr())
    torch.cuda.synchronize()
    times = []
    t_begin = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ref.solve_cuda(k, b.data_ptr(), x.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        if time.time() - t_begin > args.ref_budget_s:
            break
    ms = float(np.mean(times))
    A = to_csc(p)
    xs = x[:, :min(k, 2)].double().cpu().numpy()
    bs = b[:, :min(k, 2)].double().cpu().numpy()
    resid = float((np.linalg.norm(A @ xs - bs, axis=0) / np.linalg.norm(bs, axis=0)).max())
    line = {"impl": "reference-gpu", "metric": METRIC, "value": k / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": len(times),
            "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "none", "vs_baseline": None,
            "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": {"workload": workload_description(name, p, k, dtype), "kernels": "reference kernels/kernels.cu:64-202 (one 128-thread CTA per row)",
                       "l2": "512 MB flush write between timed steps"},
            "relative_residual": resid, "gpu_launches": 2 * len(times)}
    print(json.dumps(line), flush=True)


def to_csc(p):
    return problems.to_scipy_csc(p)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def build_solver(args, name, p, k, dtype, rank, world, device):
    """Rank 0 factorises (host, CHOLMOD) and builds the device image; with N > 1 the image (factor
    both ways + task schedule + permutation) is replicated to every rank by NCCL broadcast."""
    import torch
    import torch.distributed as dist
    from cholespy_b200 import CholeskySolverD, CholeskySolverF, _capi, multi_gpu
    lib = _capi.load()
    cls = CholeskySolverF if dtype == np.float32 else CholeskySolverD
    code = _capi.F32 if dtype == np.float32 else _capi.F64
    handle = ctypes.c_void_p()
    if rank == 0:
        perm, cp, ri, v = get_factor(name, p)
        t = time.time()
        _capi.check(lib.chb_init(device.index))
        _capi.check(lib.chb_solver_create_from_factor(ctypes.byref(handle), code, len(perm), len(ri), cp.ctypes.data,
                                                      ri.ctypes.data, v.ctypes.data, perm.ctypes.data, device.index))
        t_build = time.time() - t
        info = cls._from_handle(ctypes.c_void_p(handle.value), len(perm), device.index)
        d = info.info()
        info._handle = None  # borrowed view only
        log("device image built in %.1f s: n=%d nnz(L)=%d levels=%d/%d tasks=%d/%d (depth %d/%d) slab=%.1f/%.1f MB" % (
            t_build, d["n"], d["nnz_l"], d["n_levels_lower"], d["n_levels_upper"], d["fwd_tasks"], d["bwd_tasks"],
            d["fwd_task_levels"], d["bwd_task_levels"], d["fwd_slab_bytes"] / 1e6, d["bwd_slab_bytes"] / 1e6))
    torch.cuda.synchronize()
    t = time.time()
    solver, words, total = multi_gpu.replicate_solver(cls, handle, code, device, rank, world, dist)
    torch.cuda.synchronize()
    if rank == 0 and world > 1:
        log("NCCL broadcast of the device image: %.1f MB in %.1f ms" % (total / 1e6, 1e3 * (time.time() - t)))
    return solver, words[0], words[1], (words[10], words[11])


def run_b200(args, name, p, k_total, dtype):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # one disjoint block of host cores per rank: the pinned staging buffers are then first-touched (and the strided
        # H2D / D2H copies driven) from different cores instead of all ranks piling onto the cores of NUMA node 0
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per] or cores))
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=device)
    from cholespy_b200 import _capi
    lib = _capi.load()

    solver, n, z, nlev = build_solver(args, name, p, k_total, dtype, rank, world, device)
    s = np.dtype(dtype).itemsize
    # RHS-column shard of this rank (columns are independent; no data-path collective)
    from cholespy_b200 import multi_gpu
    # weak (default, tier rule 5): every rank solves the config's k columns of its own (k * N columns per step in
    # total, no data-path collective); strong: the config's k columns are split over the ranks
    k_config = k_total
    if args.scaling == "weak":
        k_total = k_config * world
        cols = [k_config] * world
    else:
        cols = [c for _, c in multi_gpu.column_shards(k_total, world)]
    k = cols[rank]
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    gen = torch.Generator(device="cpu").manual_seed(45 + rank)
    b_host = torch.rand((n, max(k, 1)), dtype=tdt, generator=gen).pin_memory()
    x_host = torch.empty_like(b_host).pin_memory()
    b = b_host.to(device)
    x = torch.zeros_like(b)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if k > 0:
            solver.solve(b, x)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xff)       # evict L2 between timed steps (outside the event bracket)
        ev[i][0].record()
        step()
        ev[i][1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = float(sum(a.elapsed_time(b_) for a, b_ in ev)) / args.steps
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item())

    # end to end through the public API with pinned HOST buffers (H2D + sweeps + D2H + sync per step)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        if k > 0:
            solver.solve_host(b_host, x_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if k > 0:
            solver.solve_host(b_host, x_host)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([1e3 * (time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    # parity of what was just timed: residual against the matrix (size-independent property)
    resid = None
    if rank == 0 and k > 0:
        A = problems.to_scipy_csc(p)
        kk = min(k, 4)
        xs = x[:, :kk].double().cpu().numpy()
        bs = b[:, :kk].double().cpu().numpy()
        resid = float((np.linalg.norm(A @ xs - bs, axis=0) / np.linalg.norm(bs, axis=0)).max())
        xh = x_host[:, :kk].double().numpy()
        assert np.allclose(xh, xs, rtol=1e-4, atol=1e-6), "solve_host and solve disagree"

    if rank == 0:
        info = solver.info()
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # one step = the launches of one solve (per sweep: the local-task kernel + the persistent kernel, mirrored in the
        # backward sweep; info()["launches_per_solve"]); achieved = Bytes(k) of the step / the step's device time (this
        # rank's shard).  The per-kernel shares are in the committed launch list (profiles/).
        abytes = algorithmic_bytes(n, z, k, s)
        achieved = abytes / (dev_ms * 1e-3) / 1e9
        # dram__bytes_read + dram__bytes_write of one step (summed over its launches) from the committed ncu capture of THIS
        # build: the entry names the kernels and the grid it was taken with and is ignored (traffic = null) if either differs
        traffic, traffic_source = None, None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath) and world == 1:
            entry = json.load(open(tpath)).get("%s/k%d/%s" % (args.workload, k_total, "f32" if dtype == np.float32 else "f64"))
            kernels = "chb_front" if info.get("front_image") else "chb_sweep"
            if (isinstance(entry, dict) and entry.get("grid") == info.get("grid") and kernels in str(entry.get("kernels", ""))
                    and entry.get("launches") == info.get("launches_per_solve")):
                traffic, traffic_source = entry.get("bytes_per_step"), entry.get("source")
            elif entry is not None:
                log("profiles/ncu_traffic.json entry does not match this build (grid/kernels): roofline.traffic = null")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_src, "algorithmic_bytes_per_step": abytes,
                    "traffic_per": "step", "launches_per_step": info.get("launches_per_solve", 2),
                    "roofline_columns_per_s": k / (abytes / (peak * 1e9))}
        cpu_baseline = None
        if not args.no_cpu_baseline:
            from oracle import oracle
            from helpers import max_rel_err, max_rel_err_elementwise
            tol = 1e-5 if dtype == np.float32 else 1e-12
            ck = max(1, min(k_total, args.cpu_columns))
            bc = np.random.default_rng(45).random((n, ck)).astype(dtype)
            bt = torch.tensor(bc, device=device)
            xt = torch.zeros_like(bt)
            solver.solve(bt, xt)
            torch.cuda.synchronize()
            xg = xt.cpu().numpy()
            if oracle.reference_available():
                # the reference's own CPU path on ONE core: CholeskySolver<Float>::solve_cpu -> cholmod_solve
                # (src/cholesky_solver.cpp:434-463), single-threaded by construction
                t0 = time.perf_counter()
                ref = oracle.ReferenceSolver(p["n"], p["ii"], p["jj"], p["x"], p["fmt"], dtype, cpu=True)
                t_ctor = time.perf_counter() - t0
                t0 = time.perf_counter()
                xc = ref.solve_cpu(bc)
                dt = time.perf_counter() - t0
                ref.close()
                cpu_baseline = {"value": ck / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                                "sample": "%d of %d RHS columns, one call of the reference's solve_cpu (CHOLMOD, oracle/_ref), %.1f s "
                                          "(its constructor, CHOLMOD analyze + factorize: %.0f s, not counted)" % (ck, k_total, dt, t_ctor)}
                cpu_baseline["gpu_vs_cholmod_max_rel_err"] = max_rel_err(xg, xc)
                cpu_baseline["gpu_vs_cholmod_max_rel_err_elementwise"] = max_rel_err_elementwise(xg, xc)
                assert cpu_baseline["gpu_vs_cholmod_max_rel_err"] < tol, "GPU result differs from the reference's CHOLMOD solve: %g" % cpu_baseline["gpu_vs_cholmod_max_rel_err"]
            # the C restatement of the reference's GPU kernels (oracle/cholespy_oracle.c) as the second checker, on
            # fewer columns (it is a scalar loop); it is the timed baseline only where oracle/_ref is absent
            perm, cp, ri, v = get_factor(name, p)
            ok = ck if cpu_baseline is None else min(ck, 4)
            t0 = time.perf_counter()
            xo = oracle.solve(cp, ri, v, perm, np.ascontiguousarray(bc[:, :ok]))
            dt = time.perf_counter() - t0
            if cpu_baseline is None:
                cpu_baseline = {"value": ok / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "%d of %d RHS columns, one pass of the C oracle (oracle/cholespy_oracle.c), %.1f s" % (ok, k_total, dt)}
            err = max_rel_err(xg[:, :ok], xo)
            assert err < tol, "GPU result differs from the oracle: %g" % err
            cpu_baseline["gpu_vs_oracle_max_rel_err"] = err
            cpu_baseline["gpu_vs_oracle_max_rel_err_elementwise"] = max_rel_err_elementwise(xg[:, :ok], xo)
        line = {
            "metric": METRIC, "value": k_total / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": {"workload": workload_description(name, p, k_config, dtype) + (
                           " per GPU (weak scaling: %d columns per step over %d GPUs)" % (k_total, world) if world > 1 and args.scaling == "weak"
                           else (" sharded over %d GPUs" % world if world > 1 else "")),
                       "n": n, "nnz_l": z, "k_total": k_total,
                       "k_per_rank": cols, "levels": list(nlev), "parallelism": "rhs-columns x%d" % world,
                       "l2": "512 MB flush write between timed steps", "grid": info.get("grid")},
            "e2e": {"value": k_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * k * s,
                    "d2h_bytes_per_step": n * k * s, "ms_per_step": e2e_ms},
            "gpu_launches": solver.info().get("launches_per_solve", 2) * args.steps, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "relative_residual": resid, "host_cores": os.cpu_count(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default=os.environ.get("CHB_WORKLOAD", "cfg4"))
    ap.add_argument("--k", type=int, default=0, help="override the number of RHS columns")
    ap.add_argument("--dtype", default="", choices=["", "f", "d"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = the config's RHS columns are sharded over the GPUs (default: BASELINE.json's "
                         "'128 RHS sharded over 1/2/4/8 GPUs'); weak = every GPU solves the config's columns of its own")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-columns", type=int, default=8, help="columns of the bounded CPU-baseline sample")
    ap.add_argument("--ref-procs", type=int, default=64, help="worker processes of the reference arm (capped at the core count)")
    ap.add_argument("--ref-budget-s", type=float, default=60.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    if args.impl != "b200" and rank != 0:
        return
    t = time.time()
    _, d_dtype, d_k = problems.resolve_workload(args.workload)
    p = problems.build_problem(args.workload) if rank == 0 else None  # only rank 0 touches the matrix
    k = args.k or d_k
    dtype = np.float32 if (args.dtype or d_dtype) == "f" else np.float64
    if rank == 0:
        log("problem %s built in %.1f s (n=%d, nnz(A)=%d)" % (args.workload, time.time() - t, p["n"], len(p["x"])))
    if args.impl == "reference":
        run_reference(args, args.workload, p, k, dtype)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args, args.workload, p, k, dtype)
    else:
        run_b200(args, args.workload, p, k, dtype)


if __name__ == "__main__":
    main()
