"""numpy front-end of the CPU oracle (liboracle.so) and of the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libcholespy_ref.so")
_REF_CUBIN = os.path.join(_HERE, "_ref", "ref_kernels_sm100a.cubin")

_lib = None
_ref = None


def build():
    """Compiles liboracle.so (and oracle/_ref when the reference checkout is present)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build()
        _lib = ctypes.CDLL(_ORACLE_SO)
        _lib.chb_oracle_levels.restype = ctypes.c_int
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def csc_to_csr(colptr, rowidx, vals):
    lib = _load()
    n = len(colptr) - 1
    colptr = np.ascontiguousarray(colptr, np.int32)
    rowidx = np.ascontiguousarray(rowidx, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    rowptr = np.empty(n + 1, np.int32)
    colidx = np.empty(len(rowidx), np.int32)
    rvals = np.empty(len(rowidx), np.float64)
    lib.chb_oracle_csc_to_csr(ctypes.c_int(n), _p(colptr), _p(rowidx), _p(vals), _p(rowptr), _p(colidx), _p(rvals))
    return rowptr, colidx, rvals


def levels(rowptr, colidx, descending):
    """(level_of_row, n_levels, order) of one triangular CSR operand."""
    lib = _load()
    n = len(rowptr) - 1
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    colidx = np.ascontiguousarray(colidx, np.int32)
    lvl = np.zeros(n, np.int32)
    n_levels = lib.chb_oracle_levels(ctypes.c_int(n), _p(rowptr), _p(colidx), ctypes.c_int(1 if descending else 0), _p(lvl))
    order = np.empty(n, np.int32)
    lib.chb_oracle_level_order(ctypes.c_int(n), ctypes.c_int(n_levels), _p(lvl), _p(order))
    return lvl, n_levels, order


def solve(l_colptr, l_rowidx, l_vals, perm, b, use_fma=True):
    """x = P' L^-T L^-1 P b with the reference kernels' arithmetic, in b's dtype (float32/float64).

    L is given as CSC (diag first per column, float64 values); values are cast to b's dtype first,
    like the reference does on the host (src/cholesky_solver.cpp:255-271).
    """
    lib = _load()
    dt = b.dtype
    assert dt in (np.float32, np.float64)
    sfx = "f32" if dt == np.float32 else "f64"
    n = len(l_colptr) - 1
    k = 1 if b.ndim == 1 else b.shape[1]
    l_colptr = np.ascontiguousarray(l_colptr, np.int32)
    l_rowidx = np.ascontiguousarray(l_rowidx, np.int32)
    perm = np.ascontiguousarray(perm, np.int32)
    rowptr, colidx, rvals = csc_to_csr(l_colptr, l_rowidx, l_vals)
    lower_vals = np.ascontiguousarray(rvals.astype(dt))
    upper_vals = np.ascontiguousarray(np.asarray(l_vals, np.float64).astype(dt))
    b2 = np.ascontiguousarray(b.reshape(n, k))
    tmp = np.zeros((n, k), dt)
    x = np.zeros((n, k), dt)
    getattr(lib, "chb_oracle_lower_" + sfx)(ctypes.c_int(n), ctypes.c_int(k), _p(rowptr), _p(colidx), _p(lower_vals),
                                            _p(perm), _p(b2), _p(tmp), ctypes.c_int(1 if use_fma else 0))
    getattr(lib, "chb_oracle_upper_" + sfx)(ctypes.c_int(n), ctypes.c_int(k), _p(l_colptr), _p(l_rowidx), _p(upper_vals),
                                            _p(perm), _p(tmp), _p(x), ctypes.c_int(1 if use_fma else 0))
    return x.reshape(b.shape)


# --------------------------------------------------------------------------------------------------
# The reference itself (compiled unmodified into oracle/_ref by the Makefile).
# --------------------------------------------------------------------------------------------------
def reference_available():
    return os.path.exists(_REF_SO)


def _load_ref():
    global _ref
    if _ref is None:
        if not reference_available():
            raise RuntimeError("oracle/_ref/libcholespy_ref.so is missing (build it where /root/reference exists)")
        _ref = ctypes.CDLL(_REF_SO)
        _ref.ref_create.restype = ctypes.c_void_p
        _ref.ref_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _ref.ref_last_error.restype = ctypes.c_char_p
        _ref.ref_solve_cpu.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _ref.ref_solve_cuda.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _ref.ref_get_schedule.argtypes = [ctypes.c_void_p] * 4
        _ref.ref_destroy.argtypes = [ctypes.c_void_p]
        _ref.ref_use_cubin.argtypes = [ctypes.c_void_p]
    return _ref


class ReferenceSolver:
    """The reference's CholeskySolver<Float>, driven like its bindings drive it (src/main.cpp:47-104)."""

    FMT = {"CSC": 0, "CSR": 1, "COO": 2}

    def __init__(self, n, ii, jj, x, fmt, dtype, cpu=True, use_sm100_cubin=False):
        lib = _load_ref()
        self.dtype = np.dtype(dtype)
        self.n = n
        self.cpu = cpu
        # the reference sorts/compacts some inputs in place: hand it copies
        self._ii = np.array(ii, dtype=np.int32, copy=True)
        self._jj = np.array(jj, dtype=np.int32, copy=True)
        self._x = np.array(x, dtype=np.float64, copy=True)
        if not cpu:
            if not lib.ref_init_cuda():
                raise RuntimeError("reference CUDA backend did not initialise")
            if use_sm100_cubin:
                self._image = open(_REF_CUBIN, "rb").read()
                if not lib.ref_use_cubin(self._image):
                    raise RuntimeError("could not load reference kernels compiled for sm_100a")
        self._h = lib.ref_create(1 if self.dtype == np.float64 else 0, n, len(self._x), _p(self._ii), _p(self._jj),
                                 _p(self._x), self.FMT[fmt], 1 if cpu else 0)
        if not self._h:
            raise ValueError(lib.ref_last_error().decode())

    def solve_cpu(self, b):
        b = np.ascontiguousarray(b, self.dtype)
        x = np.zeros_like(b)
        k = 1 if b.ndim == 1 else b.shape[1]
        if _load_ref().ref_solve_cpu(self._h, k, _p(b), _p(x)):
            raise ValueError(_load_ref().ref_last_error().decode())
        return x

    def solve_cuda(self, k, b_ptr, x_ptr):
        if _load_ref().ref_solve_cuda(self._h, k, ctypes.c_void_p(b_ptr), ctypes.c_void_p(x_ptr)):
            raise ValueError(_load_ref().ref_last_error().decode())

    def schedule(self):
        lo = np.empty(self.n, np.int32)
        up = np.empty(self.n, np.int32)
        pm = np.empty(self.n, np.int32)
        rc = _load_ref().ref_get_schedule(self._h, _p(lo), _p(up), _p(pm))
        if rc:
            raise RuntimeError("reference schedule copy failed (%d)" % rc)
        return lo, up, pm

    def close(self):
        if getattr(self, "_h", None):
            _load_ref().ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------
# CPU interpreter of the product's sweep schedule (oracle/panel_emulator.cpp).
# --------------------------------------------------------------------------------------------------
_EMU_SO = os.path.join(_HERE, "libpanel_emu.so")
_emu = None


def _load_emu():
    global _emu
    if _emu is None:
        if not os.path.exists(_EMU_SO):
            build()
        _emu = ctypes.CDLL(_EMU_SO)
    return _emu


SCHEDULE_INFO_KEYS = ("n_tasks", "slab_bytes", "page_bytes", "n_slots", "n_levels", "n_values", "n_stored_values",
                      "n_blocks", "n_panels", "max_local_rows")


def schedule_solve(lib, l_colptr, l_rowidx, l_vals, perm, b, exact_order=-1):
    """Builds the product's sweep schedule on the host (C ABI: chb_schedule_create) and executes it with
    the sequential CPU interpreter.  Returns (x, info_forward, info_backward)."""
    emu = _load_emu()
    dt = b.dtype
    assert dt in (np.float32, np.float64)
    n = len(l_colptr) - 1
    k = 1 if b.ndim == 1 else b.shape[1]
    l_colptr = np.ascontiguousarray(l_colptr, np.int32)
    l_rowidx = np.ascontiguousarray(l_rowidx, np.int32)
    l_vals = np.ascontiguousarray(l_vals, np.float64)
    perm = np.ascontiguousarray(perm, np.int32)
    h = ctypes.c_void_p()
    rc = lib.chb_schedule_create(ctypes.byref(h), 0 if dt == np.float32 else 1, n, len(l_rowidx), _p(l_colptr), _p(l_rowidx),
                                 _p(l_vals), _p(perm), exact_order)
    if rc:
        raise RuntimeError(lib.chb_last_error().decode())
    try:
        b2 = np.ascontiguousarray(b.reshape(n, k))
        tmp = np.zeros((n, k), dt)
        x = np.zeros((n, k), dt)
        infos = []
        fn = emu.chb_emu_sweep_f32 if dt == np.float32 else emu.chb_emu_sweep_f64
        for upper in (0, 1):
            slab, tasks, order = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            info = (ctypes.c_int64 * 10)()
            lib.chb_schedule_sweep(h, upper, ctypes.byref(slab), ctypes.byref(tasks), ctypes.byref(order), info, 10)
            info = dict(zip(SCHEDULE_INFO_KEYS, list(info)))
            infos.append(info)
            scratch = np.zeros((max(info["n_slots"], 1), 16, k), dt)
            rc = fn(slab, tasks, order, ctypes.c_int(info["n_tasks"]), ctypes.c_int(info["page_bytes"]), ctypes.c_int(upper),
                    ctypes.c_int(k), _p(tmp), _p(b2), _p(x), _p(scratch))
            if rc:
                raise RuntimeError("schedule interpreter failed with code %d (%s sweep)" % (rc, "backward" if upper else "forward"))
    finally:
        lib.chb_schedule_destroy(h)
    return x.reshape(b.shape), infos[0], infos[1]


# --------------------------------------------------------------------------------------------------
# CPU interpreter of the product's FRONT schedule (oracle/front_emulator.cpp).
# --------------------------------------------------------------------------------------------------
_FRONT_EMU_SO = os.path.join(_HERE, "libfront_emu.so")
_front_emu = None

FRONT_INFO_KEYS = ("n_tasks", "slab_bytes", "page_bytes", "tile_rows", "n_w_rows", "max_stage_rows", "n_levels", "n_values",
                   "n_stored_values", "n_local_tasks", "n_dense_tasks", "n_mma", "n_asm_entries", "n_supernodes", "n_locals",
                   "n_blocks", "n_warps", "n_local_levels", "n_local_rows")


def _load_front_emu():
    global _front_emu
    if _front_emu is None:
        if not os.path.exists(_FRONT_EMU_SO):
            build()
        _front_emu = ctypes.CDLL(_FRONT_EMU_SO)
    return _front_emu


def front_schedule_solve(lib, l_colptr, l_rowidx, l_vals, perm, b):
    """Builds the front schedule on the host (C ABI: chb_front_schedule_create) and executes both sweeps with the
    sequential CPU interpreter.  Returns (x, info_forward, info_backward)."""
    emu = _load_front_emu()
    dt = b.dtype
    assert dt in (np.float32, np.float64)
    n = len(l_colptr) - 1
    k = 1 if b.ndim == 1 else b.shape[1]
    l_colptr = np.ascontiguousarray(l_colptr, np.int32)
    l_rowidx = np.ascontiguousarray(l_rowidx, np.int32)
    l_vals = np.ascontiguousarray(l_vals, np.float64)
    perm = np.ascontiguousarray(perm, np.int32)
    h = ctypes.c_void_p()
    rc = lib.chb_front_schedule_create(ctypes.byref(h), 0 if dt == np.float32 else 1, n, len(l_rowidx), _p(l_colptr),
                                       _p(l_rowidx), _p(l_vals), _p(perm))
    if rc:
        raise RuntimeError(lib.chb_last_error().decode())
    try:
        b2 = np.ascontiguousarray(b.reshape(n, k))
        y = np.zeros((n, k), dt)
        x = np.zeros((n, k), dt)
        infos = []
        fn = emu.chb_front_emu_sweep_f32 if dt == np.float32 else emu.chb_front_emu_sweep_f64
        for upper in (0, 1):
            slab, tasks, order = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            info = (ctypes.c_int64 * 19)()
            lib.chb_front_schedule_sweep(h, upper, ctypes.byref(slab), ctypes.byref(tasks), ctypes.byref(order), info, 19)
            info = dict(zip(FRONT_INFO_KEYS, list(info)))
            infos.append(info)
            w = np.zeros((max(info["n_w_rows"], 1), k), dt)
            rc = fn(slab, tasks, order, ctypes.c_int(info["n_tasks"]), ctypes.c_int(info["page_bytes"]),
                    ctypes.c_int(info["tile_rows"]), ctypes.c_int(info["n_warps"]), ctypes.c_int(upper), ctypes.c_int(k), ctypes.c_int(n),
                    ctypes.c_int64(info["n_w_rows"]), _p(y), _p(w), _p(b2), _p(x))
            if rc:
                raise RuntimeError("front schedule interpreter failed with code %d (%s sweep)" % (rc, "backward" if upper else "forward"))
    finally:
        lib.chb_front_schedule_destroy(h)
    return x.reshape(b.shape), infos[0], infos[1]
