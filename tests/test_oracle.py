"""CPU tests: the oracle against the reference's golden vectors, and the fixtures themselves."""
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spl

import problems
from helpers import GOLDEN_CASES, etree_levels, factorize, load_golden, max_rel_err


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_solve_matches_reference_golden(case, oracle_mod):
    g = load_golden(case)
    x = oracle_mod.solve(g["l_colptr"], g["l_rowidx"], g["l_vals"], g["perm"], g["b"])
    assert max_rel_err(x, g["x_ref_cpu"]) < 1e-12
    for use_fma in (True, False):
        x32 = oracle_mod.solve(g["l_colptr"], g["l_rowidx"], g["l_vals"], g["perm"], g["b"].astype(np.float32), use_fma)
        assert x32.dtype == np.float32
        assert max_rel_err(x32, g["x_ref_cpu"]) < 1e-5
    # single right-hand side given as a 1-D array
    x1 = oracle_mod.solve(g["l_colptr"], g["l_rowidx"], g["l_vals"], g["perm"], g["b"][:, 0].copy())
    assert np.array_equal(x1, x[:, 0])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_levels_match_golden_and_etree(case, oracle_mod):
    g = load_golden(case)
    rp, ci, _ = oracle_mod.csc_to_csr(g["l_colptr"], g["l_rowidx"], g["l_vals"])
    lo, lo_n, lo_order = oracle_mod.levels(rp, ci, False)
    up, up_n, up_order = oracle_mod.levels(g["l_colptr"], g["l_rowidx"], True)
    assert np.array_equal(lo, g["lower_level"]) and np.array_equal(lo_order, g["lower_order"])
    assert np.array_equal(up, g["upper_level"]) and np.array_equal(up_order, g["upper_order"])
    assert [lo_n, up_n] == list(g["n_levels"])
    height, depth = etree_levels(g["l_colptr"], g["l_rowidx"])
    assert np.array_equal(lo, height) and np.array_equal(up, depth)
    # structure of `order`: a permutation, sorted by (level, row)
    for lvl, order in ((lo, lo_order), (up, up_order)):
        assert np.array_equal(np.sort(order), np.arange(len(order)))
        key = lvl[order].astype(np.int64) * len(order) + order
        assert np.all(np.diff(key) > 0)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_factorisation_reproduces_golden_permutation(case, native_lib):
    g = load_golden(case)
    perm, cp, ri, v = factorize(native_lib, int(g["n"]), g["ii"], g["jj"], g["x"], str(g["fmt"]))
    assert np.array_equal(perm, g["perm"])
    assert np.array_equal(cp, g["l_colptr"]) and np.array_equal(ri, g["l_rowidx"])
    assert np.array_equal(v, g["l_vals"])
    # diagonal first in every column, rows sorted
    for c in range(len(cp) - 1):
        rows = ri[cp[c]:cp[c + 1]]
        assert rows[0] == c and np.all(np.diff(rows) > 0)


def test_oracle_against_compiled_reference_cpu(oracle_mod):
    """The reference's own solve_cpu (compiled unmodified into oracle/_ref), when it is present."""
    if not oracle_mod.reference_available():
        pytest.skip("oracle/_ref not built on this box")
    p = problems.build_problem("ico4")
    rng = np.random.default_rng(3)
    b = rng.random((p["n"], 9))
    ref = oracle_mod.ReferenceSolver(p["n"], p["ii"], p["jj"], p["x"], "COO", np.float64, cpu=True)
    x_ref = ref.solve_cpu(b)
    from cholespy_b200 import _capi
    perm, cp, ri, v = factorize(_capi.load(), p["n"], p["ii"], p["jj"], p["x"], "COO")
    x = oracle_mod.solve(cp, ri, v, perm, b)
    assert max_rel_err(x, x_ref) < 1e-12
    A = problems.to_scipy_csc(p)
    assert max_rel_err(x, spl.spsolve(A, b)) < 1e-12


def test_problem_generators():
    n, f = problems.icosphere(3)
    assert n == 10 * 4 ** 3 + 2 and f.shape == (20 * 4 ** 3, 3)
    vals, idx = problems.mesh_laplacian_coo(n, f, 2.0)
    A = problems.to_scipy_csc(dict(n=n, ii=idx[0], jj=idx[1], x=vals, fmt="COO"))
    assert abs(A - A.T).max() == 0
    assert np.allclose(A @ np.ones(n), 1.0)  # I + lambda * Laplacian: rows sum to 1
    assert len(vals) > A.nnz  # duplicated diagonal entries, as in the reference's fixtures
    g = problems.grid_laplacian_2d(8).tocsc()
    assert g.shape == (64, 64) and abs(g - g.T).max() == 0
    a = problems.local_random_spd(12)
    assert abs(a - a.T).max() < 1e-9 and a.shape == (144, 144)
