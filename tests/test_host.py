"""CPU tests of the host side: the C ABI surface, matrix intake, the API mirror and its errors."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve

import problems
from cholespy_b200 import CholeskySolverD, CholeskySolverF, MatrixType, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RHS = [1, 4, 16, 64, 128, 256]


def test_c_abi_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "cholespy_b200.h")).read()
    declared = set(re.findall(r"\b(chb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(native_lib, name), "libcholespy_b200.so does not export " + name
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    assert native_lib.chb_version() >= 100


def test_gpu_entry_points_fail_loudly_without_a_device(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = problems.build_problem("cube")
    h = ctypes.c_void_p()
    rc = native_lib.chb_solver_create(ctypes.byref(h), _capi.F32, p["n"], len(p["x"]), p["ii"].ctypes.data,
                                      p["jj"].ctypes.data, p["x"].ctypes.data, 2, 0)
    assert rc == _capi.ERR_RUNTIME and not h.value
    assert b"CUDA" in native_lib.chb_last_error()


# Mirrors reference tests/test_cholesky.py:10-44 for the CPU device.
@pytest.mark.parametrize("mesh", ["cube", "ico3", "ico5"])
@pytest.mark.parametrize("variant", ["float", "double"])
def test_solver_cpu(variant, mesh):
    np.random.seed(45)
    Solver = CholeskySolverF if variant == "float" else CholeskySolverD
    dtype = np.float32 if variant == "float" else np.float64
    n_verts, faces = problems.cube() if mesh == "cube" else problems.icosphere(int(mesh[3:]))
    values, idx = problems.mesh_laplacian_coo(n_verts, faces, 2.0)
    L_csc = sp.csc_matrix((values, idx))
    b = np.random.random(size=(n_verts, max(RHS))).astype(dtype)
    x = spsolve(L_csc, b)
    solver = Solver(n_verts, idx[0], idx[1], values, MatrixType.COO)  # int64 indices: implicit conversion
    for n_rhs in RHS:
        b_crop = np.ascontiguousarray(b[:, :n_rhs])
        x_out = np.zeros_like(b_crop)
        solver.solve(b_crop, x_out)  # the CPU path has no RHS cap (reference: 256 must succeed)
        assert np.allclose(x_out, x[:, :n_rhs])
    b1 = np.ascontiguousarray(b[:, 0])
    x1 = np.zeros_like(b1)
    solver.solve(b1, x1)
    assert np.allclose(x1, x[:, 0])


# Mirrors reference tests/test_cholesky.py:46-86 for the CPU device.
def test_matrices_cpu():
    np.random.seed(45)
    n_verts, faces = problems.cube()
    values, idx = problems.mesh_laplacian_coo(n_verts, faces, 2.0)
    L_csc = sp.csc_matrix((values, idx))
    L_csr = sp.csr_matrix((values, idx))
    b = np.random.random(size=(n_verts, 32)).astype(np.float32)
    x_ref = spsolve(L_csc, b)
    for args in [(idx[0], idx[1], values, MatrixType.COO), (L_csr.indptr, L_csr.indices, L_csr.data, MatrixType.CSR),
                 (L_csc.indptr, L_csc.indices, L_csc.data, MatrixType.CSC)]:
        solver = CholeskySolverF(n_verts, *args)
        x = np.zeros_like(b)
        solver.solve(b, x)
        assert np.allclose(x, x_ref)


def test_torch_cpu_tensors():
    import torch
    np.random.seed(45)
    n_verts, faces = problems.cube()
    values, idx = problems.mesh_laplacian_coo(n_verts, faces, 2.0)
    b = np.random.random(size=(n_verts, 32)).astype(np.float32)
    x_ref = spsolve(sp.csc_matrix((values, idx)), b)
    solver = CholeskySolverF(n_verts, torch.tensor(idx[0]), torch.tensor(idx[1]), torch.tensor(values), MatrixType.COO)
    bt = torch.tensor(b)
    xt = torch.zeros_like(bt)
    solver.solve(bt, xt)
    assert np.allclose(xt.numpy(), x_ref)


def test_validation_messages():
    n_verts, faces = problems.cube()
    values, idx = problems.mesh_laplacian_coo(n_verts, faces, 2.0)
    i32 = idx.astype(np.int32)
    with pytest.raises(ValueError, match="Sparse COO matrix: the two index arrays should have the same size."):
        CholeskySolverF(n_verts, i32[0][:-1], i32[1], values, MatrixType.COO)
    with pytest.raises(ValueError, match="Sparse COO matrix: the index and data arrays should have the same size."):
        CholeskySolverF(n_verts, i32[0], i32[1], values[:-1], MatrixType.COO)
    A = sp.csr_matrix((values, idx))
    with pytest.raises(ValueError, match="Sparse CSR matrix: Invalid size for row pointer array."):
        CholeskySolverF(n_verts, A.indptr[:-1], A.indices, A.data, MatrixType.CSR)
    with pytest.raises(ValueError, match="Sparse CSC matrix: the row index and data arrays should have the same size."):
        CholeskySolverF(n_verts, A.indptr, A.indices[:-1], A.data, MatrixType.CSC)
    solver = CholeskySolverF(n_verts, i32[0], i32[1], values, MatrixType.COO)
    b = np.zeros((n_verts, 2), np.float32)
    with pytest.raises(ValueError, match="x and b should have the same dimensions."):
        solver.solve(b, np.zeros((n_verts, 3), np.float32))
    with pytest.raises(ValueError, match="Expected 1D or 2D tensors as input."):
        solver.solve(np.zeros((n_verts, 2, 2), np.float32), np.zeros((n_verts, 2, 2), np.float32))
    with pytest.raises(TypeError):  # noconvert: no implicit float64 -> float32
        solver.solve(np.zeros((n_verts, 2)), np.zeros((n_verts, 2)))
    with pytest.raises(TypeError):  # not C-contiguous
        solver.solve(np.zeros((2, n_verts), np.float32).T, np.zeros((2, n_verts), np.float32).T)


def test_not_positive_definite():
    n = 4
    ii = np.array([0, 1, 2, 3, 0, 1], np.int32)
    jj = np.array([0, 1, 2, 3, 1, 0], np.int32)
    x = np.array([1.0, 1.0, 1.0, -1.0, 3.0, 3.0])
    with pytest.raises(ValueError, match="Matrix is not positive definite!"):
        CholeskySolverD(n, ii, jj, x, MatrixType.COO)


def test_intake_sums_duplicates_and_orders_entries(native_lib):
    # unsorted COO with duplicates everywhere vs scipy's canonical form
    rng = np.random.default_rng(0)
    n = 40
    B = sp.random(n, n, density=0.15, random_state=1, format="coo")
    A = (B @ B.T + n * sp.identity(n)).tocoo()
    reps = 3
    ii = np.tile(A.row, reps).astype(np.int32)
    jj = np.tile(A.col, reps).astype(np.int32)
    xx = np.tile(A.data / reps, reps)
    shuffle = rng.permutation(len(xx))
    solver = CholeskySolverD(n, ii[shuffle], jj[shuffle], xx[shuffle], MatrixType.COO)
    b = rng.random((n, 3))
    x = np.zeros_like(b)
    solver.solve(b, x)
    assert np.allclose(x, spsolve(A.tocsc(), b), rtol=1e-10, atol=1e-12)


def test_malformed_pointer_arrays_are_rejected(native_lib):
    """A CSR/CSC pointer array that is not monotone or does not end at nnz must raise, not index out of bounds."""
    n_verts, faces = problems.cube()
    values, idx = problems.mesh_laplacian_coo(n_verts, faces, 2.0)
    A = sp.csc_matrix((values, idx))
    bad = A.indptr.astype(np.int32).copy()
    bad[3] = bad[5] + 1
    with pytest.raises(ValueError, match="invalid pointer array"):
        CholeskySolverD(n_verts, bad, A.indices.astype(np.int32), A.data, MatrixType.CSC)
    bad = A.indptr.astype(np.int32).copy()
    bad[0] = 1
    with pytest.raises(ValueError, match="invalid pointer array"):
        CholeskySolverD(n_verts, bad, A.indices.astype(np.int32), A.data, MatrixType.CSR)


def test_factor_disk_cache(native_lib, tmp_path, monkeypatch):
    """CHB_CACHE_DIR: the second factorisation of the same matrix is read back from disk (same bytes), a different
    matrix gets its own entry, and without the variable nothing is written."""
    from helpers import factorize
    p = problems.build_problem("ico3")
    first = factorize(native_lib, p["n"], p["ii"], p["jj"], p["x"], p["fmt"])
    assert not list(tmp_path.iterdir())
    monkeypatch.setenv("CHB_CACHE_DIR", str(tmp_path))
    a = factorize(native_lib, p["n"], p["ii"], p["jj"], p["x"], p["fmt"])
    files = sorted(f.name for f in tmp_path.iterdir())
    assert len(files) == 1 and files[0].startswith("factor_")
    b = factorize(native_lib, p["n"], p["ii"], p["jj"], p["x"], p["fmt"])   # cache hit
    for u, v, w in zip(first, a, b):
        assert np.array_equal(u, v) and np.array_equal(u, w)
    # a corrupted entry is ignored (and replaced), never trusted
    path = tmp_path / files[0]
    path.write_bytes(path.read_bytes()[:100])
    c = factorize(native_lib, p["n"], p["ii"], p["jj"], p["x"], p["fmt"])
    for u, v in zip(first, c):
        assert np.array_equal(u, v)
    x2 = p["x"].copy()
    x2[-1] += 1.0
    factorize(native_lib, p["n"], p["ii"], p["jj"], x2, p["fmt"])
    assert len(list(tmp_path.iterdir())) == 2
