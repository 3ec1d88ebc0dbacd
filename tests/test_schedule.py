"""CPU tests of the sweep schedule the sm_100a kernels stream (csrc/panel_schedule.cpp): the image is
built through the C ABI (chb_schedule_create, no GPU) and executed by the sequential interpreter
oracle/panel_emulator.cpp, then compared with the row-by-row oracle (cholespy_oracle.c, which follows
reference kernels/kernels.cu:81-120,149-190)."""
import numpy as np
import pytest

import problems
from helpers import TOL, factorize, max_rel_err

CASES = ["cube", "ico3", "ico5", "grid2d-40", "grid3d-10", "localspd-30"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", CASES)
def test_exact_order_schedule_is_bit_identical_to_the_oracle(name, dtype, oracle_mod, native_lib):
    p = problems.build_problem(name)
    perm, cp, ri, v = factorize(native_lib, p["n"], p["ii"], p["jj"], p["x"], p["fmt"])
    rng = np.random.default_rng(3)
    for k in (1, 5, 33):
        b = rng.random((p["n"], k)).astype(dtype)
        x_ref = oracle_mod.solve(cp, ri, v, perm, b).astype(dtype)
        x, fwd, bwd = oracle_mod.schedule_solve(native_lib, cp, ri, v, perm, b, exact_order=1)
        assert fwd["n_values"] == len(ri) and bwd["n_values"] == len(ri)  # every factor entry is used exactly once
        assert np.array_equal(x, x_ref), (name, k)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", CASES + ["ico6"])
def test_fast_schedule_within_tolerance(name, dtype, oracle_mod, native_lib):
    p = problems.build_problem(name)
    perm, cp, ri, v = factorize(native_lib, p["n"], p["ii"], p["jj"], p["x"], p["fmt"])
    b = np.random.default_rng(4).random((p["n"], 7)).astype(dtype)
    x_ref = oracle_mod.solve(cp, ri, v, perm, b)
    x, fwd, bwd = oracle_mod.schedule_solve(native_lib, cp, ri, v, perm, b, exact_order=0)
    assert fwd["n_values"] == len(ri) and bwd["n_values"] == len(ri)
    assert max_rel_err(x, x_ref) < TOL[np.dtype(dtype)], name


# Inverse super-blocks (panel_schedule.h): runs of inv_panels panels of one supernode applied through an explicit inverse,
# the inverse-apply tasks split into chunk tasks of inv_chunk blocks that park partial tiles in scratch slots.  Small
# settings make both mechanisms fire on the small test problems (the defaults, 256 / 8, need separators of >= 256 columns).
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("settings", [("4", "2"), ("3", "1"), ("64", "0"), ("1", "8")])
@pytest.mark.parametrize("name", ["ico5", "grid3d-10", "grid2d-40", "localspd-30"])
def test_inverse_super_block_settings(name, settings, dtype, oracle_mod, native_lib, monkeypatch):
    monkeypatch.setenv("CHB_INV_PANELS", settings[0])
    monkeypatch.setenv("CHB_INV_CHUNK", settings[1])
    p = problems.build_problem(name)
    perm, cp, ri, v = factorize(native_lib, p["n"], p["ii"], p["jj"], p["x"], p["fmt"])
    b = np.random.default_rng(6).random((p["n"], 5)).astype(dtype)
    x_ref = oracle_mod.solve(cp, ri, v, perm, b)
    x, fwd, bwd = oracle_mod.schedule_solve(native_lib, cp, ri, v, perm, b, exact_order=0)
    assert fwd["n_values"] == len(ri) and bwd["n_values"] == len(ri)  # every factor entry is used exactly once
    assert max_rel_err(x, x_ref) < TOL[np.dtype(dtype)], (name, settings)


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
def test_random_patterns_exact_and_fast(seed, oracle_mod, native_lib):
    """Irregular sparsity patterns (not meshes): the exact-order schedule stays bit-identical to the oracle and the
    fast schedule (relaxed panels, sparse-row local tasks, contributions, inverse super-blocks) stays within tolerance."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(20, 500))
    A = _random_spd(n, float(rng.uniform(0.002, 0.03)), seed)
    perm, cp, ri, v = factorize(native_lib, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, "CSC")
    for dtype in (np.float32, np.float64):
        b = rng.random((n, int(rng.integers(1, 9)))).astype(dtype)
        x_ref = oracle_mod.solve(cp, ri, v, perm, b)
        x_exact, fwd, bwd = oracle_mod.schedule_solve(native_lib, cp, ri, v, perm, b, exact_order=1)
        assert np.array_equal(x_exact, x_ref.astype(dtype)), (seed, n)
        assert fwd["n_values"] == len(ri) and bwd["n_values"] == len(ri)
        x_fast, fwd, bwd = oracle_mod.schedule_solve(native_lib, cp, ri, v, perm, b, exact_order=0)
        assert fwd["n_values"] == len(ri) and bwd["n_values"] == len(ri)
        assert max_rel_err(x_fast, x_ref) < TOL[np.dtype(dtype)], (seed, n)
