"""ctypes binding of include/cholespy_b200.h (libcholespy_b200.so, built in-tree by build.py)."""
import ctypes
import os

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
_LIB_PATH = os.path.join(_LIB_DIR, "libcholespy_b200%s.so" % (("_" + os.environ["CHB_LIB_VARIANT"]) if os.environ.get("CHB_LIB_VARIANT") else ""))

OK, ERR_INVALID, ERR_RUNTIME = 0, 1, 2
F32, F64 = 0, 1
DEVICE_CPU = -1
IMAGE_SHAPE_WORDS = 24

_lib = None

c_i32p = ctypes.c_void_p
c_vp = ctypes.c_void_p

_SIGNATURES = {
    "chb_last_error": (ctypes.c_char_p, []),
    "chb_version": (ctypes.c_int, []),
    "chb_init": (ctypes.c_int, [ctypes.c_int]),
    "chb_shutdown": (None, []),
    "chb_device_of_pointer": (ctypes.c_int, [c_vp]),
    "chb_sync": (ctypes.c_int, [ctypes.c_int]),
    "chb_solver_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp,
                                         ctypes.c_int, ctypes.c_int]),
    "chb_solver_create_from_device_arrays": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                            ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int]),
    "chb_solver_create_from_factor": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_vp,
                                                     c_vp, c_vp, c_vp, ctypes.c_int]),
    "chb_solver_image_shape": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "chb_solver_create_replica": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                                 ctypes.c_int, ctypes.c_int]),
    "chb_solver_destroy": (None, [c_vp]),
    "chb_solver_solve_cuda": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp]),
    "chb_solver_solve_host": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp]),
    "chb_solver_solve_cpu": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp]),
    "chb_solver_check_error": (ctypes.c_int, [c_vp]),
    "chb_solver_shard": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp]),
    "chb_solver_shard_info": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "chb_solver_is_cpu": (ctypes.c_int, [c_vp]),
    "chb_solver_info": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "chb_solver_get_analysis": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_vp]),
    "chb_solver_get_perm": (ctypes.c_int, [c_vp, c_vp]),
    "chb_solver_get_trace": (ctypes.c_int64, [c_vp, ctypes.c_int, c_vp, ctypes.c_int64]),
    "chb_factorize": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.c_int]),
    "chb_factor_n": (ctypes.c_int, [c_vp]),
    "chb_factor_nnz": (ctypes.c_int64, [c_vp]),
    "chb_factor_export": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "chb_factor_destroy": (None, [c_vp]),
    "chb_solver_image_count": (ctypes.c_int, [c_vp]),
    "chb_solver_image_buffer": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_vp),
                                               ctypes.POINTER(ctypes.c_uint64)]),
    "chb_schedule_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_vp, c_vp, c_vp,
                                           c_vp, ctypes.c_int]),
    "chb_schedule_sweep": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                                          ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "chb_schedule_destroy": (None, [c_vp]),
    "chb_front_schedule_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_vp, c_vp,
                                                 c_vp, c_vp]),
    "chb_front_schedule_sweep": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                                                ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
    "chb_front_schedule_destroy": (None, [c_vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path():
    return _LIB_PATH


def load():
    """Loads the native library; fails loudly if it has not been built (no Python fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError("cholespy_b200: native library %s is missing; run `python -m cholespy_b200.build` "
                              "(there is no pure-Python or CPU fallback for the GPU path)" % _LIB_PATH)
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc == OK:
        return
    msg = load().chb_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)
