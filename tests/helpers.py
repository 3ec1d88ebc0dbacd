"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["cube", "ico2", "ico3", "grid2d-12"]
FMT_CODE = {"CSC": 0, "CSR": 1, "COO": 2}

# Tolerances stated by BASELINE.json's north_star: max relative error vs the reference's solves.
TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: g[k] for k in g.files}


def factorize(lib, n, ii, jj, x, fmt):
    """(perm, l_colptr, l_rowidx, l_vals) through the C ABI's chb_factorize."""
    from cholespy_b200 import _capi
    h = ctypes.c_void_p()
    ii = np.ascontiguousarray(ii, np.int32)
    jj = np.ascontiguousarray(jj, np.int32)
    x = np.ascontiguousarray(x, np.float64)
    _capi.check(lib.chb_factorize(ctypes.byref(h), n, len(x), ii.ctypes.data, jj.ctypes.data, x.ctypes.data, FMT_CODE[fmt]))
    try:
        n, z = lib.chb_factor_n(h), lib.chb_factor_nnz(h)
        perm, cp = np.empty(n, np.int32), np.empty(n + 1, np.int32)
        ri, v = np.empty(z, np.int32), np.empty(z, np.float64)
        lib.chb_factor_export(h, perm.ctypes.data, cp.ctypes.data, ri.ctypes.data, v.ctypes.data)
    finally:
        lib.chb_factor_destroy(h)
    return perm, cp, ri, v


def max_rel_err(x, ref):
    """Column-wise norm-wise error max|x - ref| / max|ref| (the metric SURVEY.md 7.3-3 defines)."""
    x = np.asarray(x, np.float64).reshape(len(ref), -1)
    ref = np.asarray(ref, np.float64).reshape(len(ref), -1)
    return float((np.abs(x - ref).max(axis=0) / np.abs(ref).max(axis=0)).max())


def max_rel_err_elementwise(x, ref, floor=1e-3):
    """Element-wise error max|x - ref| / max(|ref|, floor * max|ref| of the column) (SURVEY.md 7.3-3 asks for
    both metrics; the floor keeps entries that are ~0 by cancellation from dominating)."""
    x = np.asarray(x, np.float64).reshape(len(ref), -1)
    ref = np.asarray(ref, np.float64).reshape(len(ref), -1)
    den = np.maximum(np.abs(ref), floor * np.abs(ref).max(axis=0, keepdims=True))
    return float((np.abs(x - ref) / den).max())


def rel_residual(A, x, b):
    x = np.asarray(x, np.float64).reshape(A.shape[0], -1)
    b = np.asarray(b, np.float64).reshape(A.shape[0], -1)
    return float((np.linalg.norm(A @ x - b, axis=0) / np.linalg.norm(b, axis=0)).max())


def etree_levels(l_colptr, l_rowidx):
    """Independent derivation of both level arrays from the elimination tree: for a Cholesky factor
    the forward level of a row is its height in the etree and the backward level is its depth."""
    n = len(l_colptr) - 1
    cnt = np.diff(l_colptr)
    parent = np.full(n, -1, np.int64)
    has = cnt > 1
    parent[has] = l_rowidx[l_colptr[:-1][has] + 1]
    height = np.zeros(n, np.int32)
    for j in range(n):
        p = parent[j]
        if p >= 0 and height[p] < height[j] + 1:
            height[p] = height[j] + 1
    depth = np.zeros(n, np.int32)
    for j in range(n - 1, -1, -1):
        p = parent[j]
        if p >= 0:
            depth[j] = depth[p] + 1
    return height, depth
