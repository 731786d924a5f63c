"""Deterministic synthetic matrices of the BASELINE.json configs (CRS, f64/int32/int64-ptr).

Workload generators only -- no reference code involved.  Specs: SURVEY.md section 8(d).
All matrices come back as ``(n, indptr[int64], indices[int32], vals[float64])`` with
column indices ascending within a row (what hif::CRS expects when ``check`` is on,
reference src/hif/ds/CompressedStorage.hpp:827-).
"""
from __future__ import annotations

import numpy as np


def _csr_from_coo(n, rows, cols, vals):
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    np.cumsum(indptr, out=indptr)
    return n, indptr, cols.astype(np.int32), vals.astype(np.float64)


def _stencil7(N, diag_fn, off_lo, off_hi):
    """7-point stencil on an N^3 grid, id = i + N*(j + N*k); Dirichlet = truncated rows.

    off_lo[d] / off_hi[d] are the couplings to the -/+ neighbour along axis d."""
    idx = np.arange(N ** 3, dtype=np.int64)
    i = idx % N
    j = (idx // N) % N
    k = idx // (N * N)
    coord = (i, j, k)
    stride = (1, N, N * N)
    rows = [idx]
    cols = [idx]
    nnb = np.zeros(N ** 3, dtype=np.int64)
    vals_off = []
    for d in range(3):
        lo = coord[d] > 0
        hi = coord[d] < N - 1
        nnb += lo.astype(np.int64) + hi.astype(np.int64)
        rows += [idx[lo], idx[hi]]
        cols += [idx[lo] - stride[d], idx[hi] + stride[d]]
        vals_off += [np.full(int(lo.sum()), off_lo[d]), np.full(int(hi.sum()), off_hi[d])]
    vals = [diag_fn(nnb)] + vals_off
    return _csr_from_coo(N ** 3, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))


def poisson3d(N):
    """Config 2: 3D Poisson 7-point, Dirichlet, stencil (-1,-1,-1,6,-1,-1,-1)."""
    return _stencil7(N, lambda nnb: np.full(nnb.shape, 6.0), (-1.0,) * 3, (-1.0,) * 3)


def neumann3d(N):
    """Config 5: pure-Neumann 3D Poisson (diag = #neighbours; singular, null space = constants)."""
    return _stencil7(N, lambda nnb: nnb.astype(np.float64), (-1.0,) * 3, (-1.0,) * 3)


def convdiff3d(N, conv=(20.0, 10.0, 5.0)):
    """Config 4: 3D convection-diffusion, central differences, diffusion 1,
    convection conv*h/2 added to the off-diagonals (nonsymmetric values, symmetric pattern)."""
    h = 1.0 / (N + 1)
    lo = tuple(-1.0 - c * h / 2 for c in conv)
    hi = tuple(-1.0 + c * h / 2 for c in conv)
    return _stencil7(N, lambda nnb: np.full(nnb.shape, 6.0), lo, hi)


def stokes2d_mac(N):
    """Config 3 stand-in: 2D Stokes on a staggered MAC grid of N x N cells.

    Unknowns [u: (N-1)*N ; v: N*(N-1) ; p: N*N],
    A = [[-Lap_h, 0, Gx], [0, -Lap_h, Gy], [Gx^T, Gy^T, 0]], Dirichlet walls,
    pressure pinned in cell (0,0) (row/column replaced by the identity).
    Symmetric indefinite; the zero-diagonal pressure rows force deferrals, i.e. the
    multilevel Schur recursion the config asks for."""
    nu, nv, npr = (N - 1) * N, N * (N - 1), N * N
    n = nu + nv + npr
    rows, cols, vals = [], [], []

    def add(r, c, v):
        rows.append(np.asarray(r, dtype=np.int64).ravel())
        cols.append(np.asarray(c, dtype=np.int64).ravel())
        vals.append(np.broadcast_to(np.asarray(v, dtype=np.float64), rows[-1].shape).copy())

    # u(i,j): face between cells (i,j),(i+1,j), i<N-1, j<N ; id = i + (N-1)*j
    iu, ju = np.meshgrid(np.arange(N - 1), np.arange(N), indexing="ij")
    uid = (iu + (N - 1) * ju)
    add(uid, uid, 4.0)
    m = iu > 0; add(uid[m], uid[m] - 1, -1.0)
    m = iu < N - 2; add(uid[m], uid[m] + 1, -1.0)
    m = ju > 0; add(uid[m], uid[m] - (N - 1), -1.0)
    m = ju < N - 1; add(uid[m], uid[m] + (N - 1), -1.0)
    # v(i,j): face between cells (i,j),(i,j+1), i<N, j<N-1 ; id = nu + i + N*j
    iv, jv = np.meshgrid(np.arange(N), np.arange(N - 1), indexing="ij")
    vid = nu + iv + N * jv
    add(vid, vid, 4.0)
    m = iv > 0; add(vid[m], vid[m] - 1, -1.0)
    m = iv < N - 1; add(vid[m], vid[m] + 1, -1.0)
    m = jv > 0; add(vid[m], vid[m] - N, -1.0)
    m = jv < N - 2; add(vid[m], vid[m] + N, -1.0)
    # p(i,j): id = nu + nv + i + N*j
    poff = nu + nv
    # Gx: (Gx p)_u(i,j) = p(i+1,j) - p(i,j)
    pr, pl = poff + (iu + 1) + N * ju, poff + iu + N * ju
    # Gy: (Gy p)_v(i,j) = p(i,j+1) - p(i,j)
    pt, pb = poff + iv + N * (jv + 1), poff + iv + N * jv
    pin = poff  # pinned pressure dof, cell (0,0)
    for face, pc, sgn in ((uid, pr, 1.0), (uid, pl, -1.0), (vid, pt, 1.0), (vid, pb, -1.0)):
        keep = pc != pin
        add(face[keep], pc[keep], sgn)      # G
        add(pc[keep], face[keep], sgn)      # G^T
    add(np.array([pin]), np.array([pin]), 1.0)
    return _csr_from_coo(n, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals))


def read_matrix_market(path):
    """Minimal MatrixMarket coordinate/array real reader (for examples/demo_inputs style
    files).  Returns CSR tuple for 'coordinate', a dense vector for 'array'."""
    with open(path) as f:
        header = f.readline().lower().split()
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        dims = [int(x) for x in line.split()]
        data = np.loadtxt(f, ndmin=2)
    if header[2] == "array":
        return data.reshape(-1)[: dims[0] * dims[1]].astype(np.float64)
    n = dims[0]
    rows = data[:, 0].astype(np.int64) - 1
    cols = data[:, 1].astype(np.int64) - 1
    vals = data[:, 2].astype(np.float64)
    if header[4] == "symmetric":
        off = rows != cols
        rows, cols, vals = (np.concatenate([rows, cols[off]]), np.concatenate([cols, rows[off]]),
                            np.concatenate([vals, vals[off]]))
    return _csr_from_coo(n, rows, cols, vals)


def csr_matvec(A, x):
    """y = A x for a CSR tuple (numpy; test helper)."""
    n, indptr, indices, vals = A
    prod = vals * x[indices]
    y = np.add.reduceat(prod, indptr[:-1].astype(np.int64)) if len(prod) else np.zeros(n)
    empty = indptr[1:] == indptr[:-1]
    if empty.any():
        y[empty] = 0.0
    return y


def seeded_rhs(n, j=0, nrhs=None):
    """Apply inputs of SURVEY.md 8(d): b_j[i] ~ U(-1,1), seed 20260101+j (numpy PCG64 stream;
    the survey's mt19937_64 stream is a C++ detail -- only determinism matters here)."""
    if nrhs is None:
        return np.random.default_rng(20260101 + j).uniform(-1.0, 1.0, n)
    return np.stack([seeded_rhs(n, j + k) for k in range(nrhs)], axis=1).copy()


PDE_PARAMS = dict(tau_L=1e-2, tau_U=1e-2, kappa_d=5.0, kappa=5.0, alpha_L=3.0, alpha_U=3.0)
"""The reference's parameter set for well-posed PDE systems
(examples/advanced/demo_gmreshif.cpp:62-66, docs params.doc:53-63)."""
