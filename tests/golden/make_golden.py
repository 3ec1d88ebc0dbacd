"""Regenerates tests/golden/*.npz.  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

For each small case it stores the inputs (COO arrays, b), the factor produced by the reference's
CHOLMOD configuration (perm, L in CSC), the level arrays of both triangles, and the solution of the
REFERENCE ITSELF: oracle/_ref/libcholespy_ref.so is the reference's unmodified src/cholesky_solver.cpp
compiled here, and x_ref_cpu is its solve_cpu() output.  Before writing, the script checks that the
reference's fixtures (tests/utils.py, imported from /root/reference) produce the same matrices as
tests/problems.py, and that the C oracle reproduces x_ref_cpu.
"""
import ctypes
import os
import sys

import numpy as np
import scipy.sparse.linalg as spl

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems  # noqa: E402
from oracle import oracle  # noqa: E402
from cholespy_b200 import _capi  # noqa: E402

CASES = ["cube", "ico2", "ico3", "grid2d-12"]
REF_TESTS = "/root/reference/tests"


def check_fixtures_against_reference():
    sys.path.insert(0, REF_TESTS)
    import utils  # the reference's own fixtures
    n, f = utils.get_cube()
    n2, f2 = problems.cube()
    assert n == n2 and np.array_equal(f, f2)
    for level in range(0, 6):
        n, f = utils.get_icosphere(level)
        n2, f2 = problems.icosphere(level)
        assert n == n2 and np.array_equal(f, f2), level
        for lam in (2.0, 10.0):
            v, i = utils.get_coo_arrays(n, f, lam)
            v2, i2 = problems.mesh_laplacian_coo(n2, f2, lam)
            assert np.array_equal(v, v2) and np.array_equal(i, i2)
    print("fixtures identical to the reference's tests/utils.py (cube, icosphere 0-5)")


def factorize(p):
    lib = _capi.load()
    h = ctypes.c_void_p()
    fmt = {"CSC": 0, "CSR": 1, "COO": 2}[p["fmt"]]
    _capi.check(lib.chb_factorize(ctypes.byref(h), p["n"], len(p["x"]), p["ii"].ctypes.data, p["jj"].ctypes.data,
                                  p["x"].ctypes.data, fmt))
    n, z = lib.chb_factor_n(h), lib.chb_factor_nnz(h)
    perm, cp = np.empty(n, np.int32), np.empty(n + 1, np.int32)
    ri, v = np.empty(z, np.int32), np.empty(z, np.float64)
    lib.chb_factor_export(h, perm.ctypes.data, cp.ctypes.data, ri.ctypes.data, v.ctypes.data)
    lib.chb_factor_destroy(h)
    return perm, cp, ri, v


def main():
    check_fixtures_against_reference()
    rng = np.random.default_rng(45)
    for name in CASES:
        p = problems.build_problem(name)
        A = problems.to_scipy_csc(p)
        perm, cp, ri, v = factorize(p)
        rp, ci, rv = oracle.csc_to_csr(cp, ri, v)
        lo_lvl, lo_n, lo_order = oracle.levels(rp, ci, False)
        up_lvl, up_n, up_order = oracle.levels(cp, ri, True)
        b = rng.random((p["n"], 5))
        ref = oracle.ReferenceSolver(p["n"], p["ii"], p["jj"], p["x"], p["fmt"], np.float64, cpu=True)
        x_ref = ref.solve_cpu(b)
        ref32 = oracle.ReferenceSolver(p["n"], p["ii"], p["jj"], p["x"], p["fmt"], np.float32, cpu=True)
        x_ref32 = ref32.solve_cpu(b.astype(np.float32))
        x_scipy = spl.spsolve(A, b)
        x_or = oracle.solve(cp, ri, v, perm, b)
        err = np.abs(x_or - x_ref).max() / np.abs(x_ref).max()
        assert err < 1e-13, (name, err)
        assert np.abs(x_ref - x_scipy).max() / np.abs(x_scipy).max() < 1e-12
        np.savez_compressed(os.path.join(HERE, name + ".npz"), n=p["n"], ii=p["ii"], jj=p["jj"], x=p["x"], fmt=p["fmt"],
                            perm=perm, l_colptr=cp, l_rowidx=ri, l_vals=v, lower_level=lo_lvl, lower_order=lo_order,
                            upper_level=up_lvl, upper_order=up_order, n_levels=np.array([lo_n, up_n]), b=b,
                            x_ref_cpu=x_ref, x_ref_cpu_f32=x_ref32)
        print("%-10s n=%d nnz(L)=%d levels=%d/%d oracle-vs-reference %.2e" % (name, p["n"], len(ri), lo_n, up_n, err))


if __name__ == "__main__":
    main()
