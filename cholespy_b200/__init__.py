"""cholespy_b200 -- B200-native drop-in for cholespy's solver classes.

Same Python surface as the reference package (reference: src/main.cpp:16-135, src/docstr.h,
src/cholespy/__init__.py): `CholeskySolverF`, `CholeskySolverD`, `MatrixType`, constructor
`(n_rows, ii, jj, x, type)` and in-place `.solve(b, x)`, same validation messages and exception
types.  The GPU path runs hand-written sm_100a kernels through the C ABI in
include/cholespy_b200.h; arrays on a CUDA device select it, host arrays select the CHOLMOD CPU
solver, exactly as in the reference.  There is no CPU fallback for CUDA inputs.
"""
import ctypes
import enum

import numpy as np

from . import _capi

__version__ = "0.1.0"
__all__ = ["CholeskySolverF", "CholeskySolverD", "MatrixType"]


class MatrixType(enum.IntEnum):
    """Sparse matrix layout of the constructor arrays (reference: src/cholesky_solver.h:8-12)."""
    CSC = 0
    CSR = 1
    COO = 2


class _Array:
    """Borrowed view of a framework array: raw pointer, shape, dtype and device."""
    __slots__ = ("ptr", "shape", "dtype", "device", "owner")

    def __init__(self, ptr, shape, dtype, device, owner):
        self.ptr, self.shape, self.dtype, self.device, self.owner = ptr, tuple(shape), np.dtype(dtype), device, owner

    @property
    def is_cuda(self):
        return self.device >= 0


def _is_torch(obj):
    return type(obj).__module__.split(".")[0] == "torch" and hasattr(obj, "data_ptr")


def _adapt(obj, dtype=None, convert=False, name="array"):
    """Returns a C-contiguous _Array for numpy / torch / __cuda_array_interface__ / DLPack inputs.

    convert=True allows implicit dtype conversion and compaction (constructor arguments, like
    nanobind's default); convert=False demands the exact dtype and C-contiguity (`solve`'s
    noconvert arguments, reference src/main.cpp:108-109) and raises TypeError otherwise.
    """
    want = None if dtype is None else np.dtype(dtype)
    if _is_torch(obj):
        import torch
        t = obj
        tdt = np.dtype(str(t.dtype).replace("torch.", ""))
        if want is not None and tdt != want:
            if not convert:
                raise TypeError("%s: expected dtype %s, got %s" % (name, want, tdt))
            t = t.to(getattr(torch, want.name))
        if not t.is_contiguous():
            if not convert:
                raise TypeError("%s: expected a C-contiguous array" % name)
            t = t.contiguous()
        t = t.detach()
        dev = t.device.index if t.device.type == "cuda" else -1
        if t.device.type == "cuda" and dev is None:
            dev = torch.cuda.current_device()
        if t.device.type not in ("cuda", "cpu"):
            raise ValueError("Unsupported input device! Only CPU and CUDA arrays are supported.")
        return _Array(t.data_ptr(), t.shape, want or tdt, dev, t)
    if hasattr(obj, "__cuda_array_interface__"):
        cai = obj.__cuda_array_interface__
        adt = np.dtype(cai["typestr"])
        shape = tuple(cai["shape"])
        strides = cai.get("strides")
        if strides is not None:
            expect, acc = [], adt.itemsize
            for s in reversed(shape):
                expect.append(acc)
                acc *= s
            if tuple(strides) != tuple(reversed(expect)):
                raise TypeError("%s: expected a C-contiguous array" % name)
        if want is not None and adt != want:
            raise TypeError("%s: expected dtype %s, got %s" % (name, want, adt))
        ptr = cai["data"][0]
        dev = _capi.load().chb_device_of_pointer(ptr) if ptr else 0
        return _Array(ptr, shape, adt, max(dev, 0), obj)
    if not isinstance(obj, np.ndarray) and hasattr(obj, "__dlpack__") and not hasattr(obj, "__array_interface__"):
        try:
            import torch
            return _adapt(torch.from_dlpack(obj), dtype, convert, name)
        except ImportError:
            obj = np.from_dlpack(obj)
    a = np.asarray(obj)
    if want is not None and a.dtype != want:
        if not convert:
            raise TypeError("%s: expected dtype %s, got %s" % (name, want, a.dtype))
        a = a.astype(want)
    if not a.flags.c_contiguous:
        if not convert:
            raise TypeError("%s: expected a C-contiguous array" % name)
        a = np.ascontiguousarray(a)
    return _Array(a.ctypes.data, a.shape, a.dtype, -1, a)


class _CholeskySolver:
    _DTYPE = None       # numpy dtype of solve() arrays
    _DTYPE_CODE = None  # CHB_F32 / CHB_F64

    def __init__(self, n_rows, ii, jj, x, type):
        self._handle = None
        lib = _capi.load()
        n_rows = int(n_rows)
        if n_rows < 0:
            raise TypeError("n_rows must be an unsigned integer")
        mtype = MatrixType(type)
        a_ii = _adapt(ii, np.int32, True, "ii")
        a_jj = _adapt(jj, np.int32, True, "jj")
        a_x = _adapt(x, np.float64, True, "x")
        for a in (a_ii, a_jj, a_x):
            if len(a.shape) != 1:
                raise TypeError("constructor arrays must be one-dimensional")
        # Same checks and messages as the reference bindings (src/main.cpp:28-45).
        if mtype == MatrixType.COO:
            if a_ii.shape[0] != a_jj.shape[0]:
                raise ValueError("Sparse COO matrix: the two index arrays should have the same size.")
            if a_ii.shape[0] != a_x.shape[0]:
                raise ValueError("Sparse COO matrix: the index and data arrays should have the same size.")
        elif mtype == MatrixType.CSR:
            if a_jj.shape[0] != a_x.shape[0]:
                raise ValueError("Sparse CSR matrix: the column index and data arrays should have the same size.")
            if a_ii.shape[0] != n_rows + 1:
                raise ValueError("Sparse CSR matrix: Invalid size for row pointer array.")
        else:
            if a_jj.shape[0] != a_x.shape[0]:
                raise ValueError("Sparse CSC matrix: the row index and data arrays should have the same size.")
            if a_ii.shape[0] != n_rows + 1:
                raise ValueError("Sparse CSC matrix: Invalid size for column pointer array.")
        if a_ii.is_cuda != a_jj.is_cuda or a_ii.is_cuda != a_x.is_cuda:
            raise ValueError("All input tensors should be on the same device!")
        handle = ctypes.c_void_p()
        nnz = a_x.shape[0]
        if a_ii.is_cuda:
            self._device = a_ii.device
            _capi.check(lib.chb_init(self._device))
            _capi.check(lib.chb_solver_create_from_device_arrays(
                ctypes.byref(handle), self._DTYPE_CODE, n_rows, a_ii.shape[0], nnz, a_ii.ptr, a_jj.ptr, a_x.ptr,
                int(mtype), self._device))
        else:
            self._device = _capi.DEVICE_CPU
            _capi.check(lib.chb_solver_create(ctypes.byref(handle), self._DTYPE_CODE, n_rows, nnz, a_ii.ptr, a_jj.ptr,
                                              a_x.ptr, int(mtype), _capi.DEVICE_CPU))
        self._handle = handle
        self._n = n_rows

    @classmethod
    def _from_handle(cls, handle, n, device):
        self = cls.__new__(cls)
        self._handle, self._n, self._device = handle, n, device
        return self

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _capi.load().chb_solver_destroy(h)
            except Exception:
                pass

    def solve(self, b, x):
        a_b = _adapt(b, self._DTYPE, False, "b")
        a_x = _adapt(x, self._DTYPE, False, "x")
        # Same checks and messages as the reference bindings (src/main.cpp:83-106).
        if len(a_b.shape) not in (1, 2):
            raise ValueError("Expected 1D or 2D tensors as input.")
        if a_b.shape[0] != a_x.shape[0] or len(a_b.shape) != len(a_x.shape) or \
                (len(a_b.shape) == 2 and a_b.shape[1] != a_x.shape[1]):
            raise ValueError("x and b should have the same dimensions.")
        if a_b.is_cuda != a_x.is_cuda:
            raise ValueError("x and b should be on the same device.")
        n_rhs = a_b.shape[1] if len(a_b.shape) == 2 else 1
        lib = _capi.load()
        cpu_solver = self._device == _capi.DEVICE_CPU
        if not a_b.is_cuda:
            if not cpu_solver:
                raise ValueError("Input device is CPU but the solver was initialized for CUDA.")
            _capi.check(lib.chb_solver_solve_cpu(self._handle, n_rhs, a_b.ptr, a_x.ptr))
        else:
            if cpu_solver:
                raise ValueError("Input device is CUDA but the solver was initialized for CPU.")
            # (new, the reference is single-device: src/cuda_driver.cpp:124) foreign pointers would fault the
            # solver's context, which it shares with the caller's framework
            for a in (a_b, a_x):
                if a.device is not None and a.device != self._device:
                    raise ValueError("x and b should be on the solver's device (cuda:%d)." % self._device)
            _capi.check(lib.chb_solver_solve_cuda(self._handle, n_rhs, a_b.ptr, a_x.ptr))
        return None

    # -- additions (not in the reference API) ----------------------------------------------------
    def solve_host(self, b, x):
        """End-to-end solve of HOST arrays on the GPU (H2D, two sweeps, D2H, sync)."""
        a_b = _adapt(b, self._DTYPE, False, "b")
        a_x = _adapt(x, self._DTYPE, False, "x")
        if a_b.is_cuda or a_x.is_cuda or a_b.shape != a_x.shape:
            raise ValueError("solve_host expects two host arrays of the same shape.")
        n_rhs = a_b.shape[1] if len(a_b.shape) == 2 else 1
        _capi.check(_capi.load().chb_solver_solve_host(self._handle, n_rhs, a_b.ptr, a_x.ptr))

    def info(self):
        out = (ctypes.c_int64 * 21)()
        m = _capi.load().chb_solver_info(self._handle, out, 21)
        keys = ("n", "nnz_l", "n_levels_lower", "n_levels_upper", "dtype", "device", "grid", "launches_per_solve",
                "fwd_tasks", "fwd_slab_bytes", "fwd_slots", "fwd_task_levels", "fwd_stored_values",
                "bwd_tasks", "bwd_slab_bytes", "bwd_slots", "bwd_task_levels", "bwd_stored_values",
                "front_image", "update_rows", "compute_warps")
        return dict(zip(keys[:m], list(out)[:m]))

    def shard(self, devices):
        """Shards the RHS columns of later solves over further GPUs (ordinals of the ADDITIONAL devices; [] undoes it).
        New: the reference is single-device.  See include/cholespy_b200.h: chb_solver_shard."""
        dev = np.ascontiguousarray(list(devices), np.int32)
        _capi.check(_capi.load().chb_solver_shard(self._handle, len(dev), dev.ctypes.data if len(dev) else None))

    def shard_info(self):
        out = (ctypes.c_int64 * 4)()
        _capi.load().chb_solver_shard_info(self._handle, out, 4)
        return {"devices": int(out[0]), "replication": {0: "none", 1: "nccl", 2: "peer"}[int(out[1])], "bytes": int(out[2]),
                "seconds": out[3] * 1e-6}

    def check_error(self):
        """Synchronises the device and raises RuntimeError if the sweep kernels' watchdog fired."""
        _capi.check(_capi.load().chb_solver_check_error(self._handle))

    def analysis(self, lower=True):
        """(level_of_row, order): the GPU level-set analysis products (int32 numpy arrays)."""
        lvl = np.empty(self._n, np.int32)
        order = np.empty(self._n, np.int32)
        _capi.check(_capi.load().chb_solver_get_analysis(self._handle, 1 if lower else 0, lvl.ctypes.data, order.ctypes.data))
        return lvl, order

    def permutation(self):
        perm = np.empty(self._n, np.int32)
        _capi.check(_capi.load().chb_solver_get_perm(self._handle, perm.ctypes.data))
        return perm


class CholeskySolverF(_CholeskySolver):
    """Single-precision Cholesky solver (reference: src/docstr.h, CholeskySolverF)."""
    _DTYPE = np.float32
    _DTYPE_CODE = _capi.F32


class CholeskySolverD(_CholeskySolver):
    """Double-precision Cholesky solver (reference: src/docstr.h, CholeskySolverD)."""
    _DTYPE = np.float64
    _DTYPE_CODE = _capi.F64
