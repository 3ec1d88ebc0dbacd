"""CPU tests of the FRONT schedule the default sm_100a sweep kernels stream (csrc/front_schedule.cpp): the image is
built through the C ABI (chb_front_schedule_create, no GPU), executed by the sequential interpreter
oracle/front_emulator.cpp (which also checks that ticket order respects every dependency and that no row is read
before it is written), and compared with the row-by-row oracle (cholespy_oracle.c, which follows reference
kernels/kernels.cu:81-120,149-190)."""
import numpy as np
import pytest

import problems
from helpers import TOL, factorize, max_rel_err, max_rel_err_elementwise

CASES = ["cube", "ico3", "ico5", "grid2d-40", "grid3d-10", "localspd-30"]


def _check(name, dtype, k, oracle_mod, native_lib, seed=4):
    p = problems.build_problem(name)
    perm, cp, ri, v = factorize(native_lib, p["n"], p["ii"], p["jj"], p["x"], p["fmt"])
    b = np.random.default_rng(seed).random((p["n"], k)).astype(dtype)
    x_ref = oracle_mod.solve(cp, ri, v, perm, b)
    x, fwd, bwd = oracle_mod.front_schedule_solve(native_lib, cp, ri, v, perm, b)
    assert fwd["n_values"] == len(ri) and bwd["n_values"] == len(ri)  # every factor entry is used exactly once per sweep
    assert max_rel_err(x, x_ref) < TOL[np.dtype(dtype)], name
    assert max_rel_err_elementwise(x, x_ref) < 4 * TOL[np.dtype(dtype)], name
    return fwd, bwd


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", CASES + ["ico6"])
def test_front_schedule_within_tolerance(name, dtype, oracle_mod, native_lib):
    _check(name, dtype, 7, oracle_mod, native_lib)


@pytest.mark.parametrize("k", [1, 3, 33, 128])
def test_front_schedule_rhs_widths(k, oracle_mod, native_lib):
    _check("ico4", np.float32, k, oracle_mod, native_lib)
    _check("grid2d-24", np.float64, k, oracle_mod, native_lib)


def test_front_schedule_wide_supernodes_split_into_panels(oracle_mod, native_lib, monkeypatch):
    """Small panels / chunk budgets force every code path of the builder on a small problem: multi-panel supernodes,
    partial-product tasks, separate below-product tasks with several column ranges."""
    monkeypatch.setenv("CHB_FRONT_PANEL", "32")
    monkeypatch.setenv("CHB_FRONT_SPLIT_CHUNKS", "1")
    monkeypatch.setenv("CHB_FRONT_OWN_CHUNKS", "1")
    monkeypatch.setenv("CHB_FRONT_FUSE_STEPS", "0")
    monkeypatch.setenv("CHB_FRONT_LOCAL_ROWS", "24")
    monkeypatch.setenv("CHB_FRONT_LOCAL_EXT", "16")
    monkeypatch.setenv("CHB_FRONT_ORDER_BLOCKS", "6")
    for name in ("ico5", "grid3d-10", "localspd-30"):
        for dtype in (np.float32, np.float64):
            fwd, bwd = _check(name, dtype, 5, oracle_mod, native_lib)
            assert fwd["n_dense_tasks"] > 0 and fwd["n_local_tasks"] > 0 and fwd["n_asm_entries"] > 0


def test_front_schedule_no_local_tasks(oracle_mod, native_lib, monkeypatch):
    monkeypatch.setenv("CHB_FRONT_LOCAL_ROWS", "1")
    monkeypatch.setenv("CHB_FRONT_LOCAL_EXT", "0")
    _check("ico3", np.float64, 4, oracle_mod, native_lib)


def _random_spd(n, density, seed):
    """A = B B' + n I with a random sparse B plus a band (irregular elimination trees: chains, stars, wide fronts)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    B = sp.random(n, n, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csr")
    band = sp.diags([rng.standard_normal(n - 1), rng.standard_normal(max(n - 7, 0))], [-1, -7], shape=(n, n))
    B = (B + band).tocsr()
    A = (B @ B.T + float(n) * sp.identity(n)).tocsc()
    A.sort_indices()
    return A


@pytest.mark.parametrize("seed", range(8))
def test_front_schedule_random_patterns(seed, oracle_mod, native_lib):
    rng = np.random.default_rng(200 + seed)
    n = int(rng.integers(20, 600))
    A = _random_spd(n, float(rng.uniform(0.002, 0.03)), seed)
    perm, cp, ri, v = factorize(native_lib, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, "CSC")
    for dtype in (np.float32, np.float64):
        b = rng.random((n, int(rng.integers(1, 9)))).astype(dtype)
        x_ref = oracle_mod.solve(cp, ri, v, perm, b)
        x, fwd, bwd = oracle_mod.front_schedule_solve(native_lib, cp, ri, v, perm, b)
        assert fwd["n_values"] == len(ri) and bwd["n_values"] == len(ri)
        assert max_rel_err(x, x_ref) < TOL[np.dtype(dtype)], (seed, n)


def test_front_schedule_degenerate(oracle_mod, native_lib):
    """1 x 1, diagonal and block-diagonal systems (forests, roots without below-rows)."""
    import scipy.sparse as sp
    for A in (sp.csc_matrix(np.array([[4.0]])), sp.identity(9, format="csc") * 3.0,
              sp.block_diag([np.array([[4.0, 1.0], [1.0, 3.0]])] * 5, format="csc")):
        A.sort_indices()
        n = A.shape[0]
        perm, cp, ri, v = factorize(native_lib, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, "CSC")
        b = np.random.default_rng(1).random((n, 3))
        x, _, _ = oracle_mod.front_schedule_solve(native_lib, cp, ri, v, perm, b)
        assert np.allclose(A @ x, b, rtol=1e-12, atol=1e-14)
