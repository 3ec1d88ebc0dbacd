"""Synthetic SPD systems for the parity tests and bench.py (BASELINE.json `configs`).

Test/bench infrastructure, not product code.  The mesh recipes reproduce, vectorised, exactly what
the reference's fixtures build (reference: tests/utils.py:3-18 `get_coo_arrays`, :23-82
`get_icosphere`, :84-97 `get_cube`): same vertex numbering, same COO entry order, same duplicated
diagonal entries, so matrices are identical to the ones the reference's own tests factorise.
`tests/golden/make_golden.py` pins that equivalence against the reference implementation.
"""
import numpy as np
import scipy.sparse as sp

_ICO_FACES = np.array([
    [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
    [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
    [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
    [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)


def icosphere(level):
    """(n_verts, faces) of the `level`-times subdivided icosahedron, n = 10*4**level + 2.

    New vertices are numbered in order of first appearance while walking faces in order and, inside
    a face, edges (0,1), (1,2), (2,0) -- the numbering the reference's sequential routine produces.
    """
    faces = _ICO_FACES.copy()
    n_verts = 12
    for _ in range(level):
        m = faces.shape[0]
        a = faces[:, [0, 1, 2]].reshape(-1)
        b = faces[:, [1, 2, 0]].reshape(-1)
        key = np.minimum(a, b) * np.int64(n_verts) + np.maximum(a, b)
        _, first, inverse = np.unique(key, return_index=True, return_inverse=True)
        rank = np.empty_like(first)
        rank[np.argsort(first, kind="stable")] = np.arange(first.size)
        mid = (n_verts + rank[inverse]).reshape(m, 3)  # mid[f] = (v1, v2, v3)
        t0, t1, t2 = faces[:, 0], faces[:, 1], faces[:, 2]
        v1, v2, v3 = mid[:, 0], mid[:, 1], mid[:, 2]
        faces = np.stack([np.stack([t0, v1, v3], 1), np.stack([t1, v2, v1], 1),
                          np.stack([t2, v3, v2], 1), np.stack([v1, v2, v3], 1)], axis=1).reshape(4 * m, 3)
        n_verts += first.size
    return int(n_verts), faces


def cube():
    faces = np.array([[0, 1, 3], [0, 3, 2], [2, 3, 7], [2, 7, 6], [4, 5, 7], [4, 7, 6],
                      [0, 1, 5], [0, 5, 4], [1, 5, 7], [1, 7, 3], [0, 4, 6], [0, 6, 2]])
    return 8, faces


def mesh_laplacian_coo(n_verts, faces, lambda_):
    """COO arrays (values, idx[2, nnz]) of I + lambda * L_uniform, WITH duplicate diagonal entries."""
    ii = faces[:, [1, 2, 0]].flatten()
    jj = faces[:, [2, 0, 1]].flatten()
    r = np.concatenate([ii, jj]).astype(np.int64)
    c = np.concatenate([jj, ii]).astype(np.int64)
    key = np.unique(r * n_verts + c)  # lexicographic (row, col) order == np.unique(axis=1) on the pair
    adj = np.stack([key // n_verts, key % n_verts], axis=0)
    adj_values = np.full(adj.shape[1], float(lambda_), dtype=np.float64)
    diag_idx = np.stack((adj[0], adj[0]), axis=0)
    diag = np.stack((np.arange(n_verts), np.arange(n_verts)), axis=0)
    idx = np.concatenate((adj, diag_idx, diag), axis=1)
    values = np.concatenate((-adj_values, adj_values, np.ones(n_verts)))
    return values, idx


def grid_laplacian_2d(m, shift=1.0):
    """5-point Laplacian on an m x m grid (Dirichlet) + shift*I as scipy COO (cfg 1: m=256)."""
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m))
    I = sp.identity(m)
    A = sp.kron(I, T) + sp.kron(T, I) + shift * sp.identity(m * m)
    return A.tocoo()


def grid_laplacian_3d(m, shift=0.0):
    """7-point Laplacian on an m^3 grid (Dirichlet, already SPD) as scipy CSC (cfg 3: m=128)."""
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m))
    I = sp.identity(m)
    A = sp.kron(sp.kron(I, I), T) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(T, I), I)
    if shift:
        A = A + shift * sp.identity(m ** 3)
    A = A.tocsc()
    A.sort_indices()
    return A


def local_random_spd(side, draws=4, radius=1, seed=7):
    """cfg 5 stand-in: A = B'B + n*I with B's non-zeros drawn inside a (2r+1)^2 neighbourhood of a
    side x side grid (a uniformly random B fills in quadratically and cannot be factorised, see
    SURVEY.md 7.3-2).  Returns scipy CSR with sorted indices."""
    rng = np.random.default_rng(seed)
    n = side * side
    rows = np.repeat(np.arange(n), draws)
    gx, gy = rows % side, rows // side
    dx = rng.integers(-radius, radius + 1, size=rows.size)
    dy = rng.integers(-radius, radius + 1, size=rows.size)
    cx = np.clip(gx + dx, 0, side - 1)
    cy = np.clip(gy + dy, 0, side - 1)
    B = sp.csr_matrix((rng.standard_normal(rows.size), (rows, cy * side + cx)), shape=(n, n))
    A = (B.T @ B + float(n) * sp.identity(n)).tocsr()
    A.sort_indices()
    return A


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs -> (name, n, (ii, jj, x, MatrixType name), dtype, n_rhs, scipy matrix)
# ------------------------------------------------------------------------------------------------
def resolve_workload(name):
    """(generator name, dtype 'f'|'d', n_rhs) of a workload name or BASELINE alias, without building it."""
    import os
    alias = {
        "cfg1": ("grid2d-256", "d", 1),
        "cfg2": ("ico8", "f", 3),
        "cfg3": ("grid3d-%d" % int(os.environ.get("CHB_CFG3_SIDE", "64")), "d", 1),
        "cfg4": ("ico9", "f", 128),
        "cfg5": ("localspd-1000", "d", 128),
    }
    return alias.get(name, (name, "f", 1))


def build_problem(name):
    """Returns dict(n, ii, jj, x, fmt, dtype, k, A) for a named workload.

    names: cube, ico<L>[f|d]-k<K>, grid2d-<m>, grid3d-<m>, localspd-<side>, and the aliases
    cfg1..cfg5 for BASELINE.json's configs (cfg3 -> 3-D grid 64^3 stand-in unless CHB_CFG3_SIDE is
    set; cfg5 -> the locality-structured stand-in).  k/dtype can be overridden by the caller.
    """
    name, dtype, k = resolve_workload(name)
    if name == "cube":
        n, faces = cube()
        vals, idx = mesh_laplacian_coo(n, faces, 2.0)
        return dict(n=n, ii=idx[0].astype(np.int32), jj=idx[1].astype(np.int32), x=vals, fmt="COO", dtype=dtype, k=k)
    if name.startswith("ico"):
        level = int(name[3:])
        n, faces = icosphere(level)
        vals, idx = mesh_laplacian_coo(n, faces, 10.0)
        return dict(n=n, ii=idx[0].astype(np.int32), jj=idx[1].astype(np.int32), x=vals, fmt="COO", dtype=dtype, k=k)
    if name.startswith("grid2d-"):
        A = grid_laplacian_2d(int(name.split("-")[1]))
        return dict(n=A.shape[0], ii=A.row.astype(np.int32), jj=A.col.astype(np.int32), x=A.data.astype(np.float64),
                    fmt="COO", dtype=dtype, k=k)
    if name.startswith("grid3d-"):
        A = grid_laplacian_3d(int(name.split("-")[1]))
        return dict(n=A.shape[0], ii=A.indptr.astype(np.int32), jj=A.indices.astype(np.int32),
                    x=A.data.astype(np.float64), fmt="CSC", dtype=dtype, k=k)
    if name.startswith("localspd-"):
        A = local_random_spd(int(name.split("-")[1]))
        return dict(n=A.shape[0], ii=A.indptr.astype(np.int32), jj=A.indices.astype(np.int32),
                    x=A.data.astype(np.float64), fmt="CSR", dtype=dtype, k=k)
    raise ValueError("unknown problem " + name)


def to_scipy_csc(p):
    if p["fmt"] == "COO":
        return sp.csc_matrix((p["x"], (p["ii"], p["jj"])), shape=(p["n"], p["n"]))
    if p["fmt"] == "CSR":
        return sp.csr_matrix((p["x"], p["jj"], p["ii"]), shape=(p["n"], p["n"])).tocsc()
    return sp.csc_matrix((p["x"], p["jj"], p["ii"]), shape=(p["n"], p["n"]))
