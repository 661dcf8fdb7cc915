"""The Krylov oracle (numpy restatement) must agree with scipy.sparse.linalg itself.  CPU only."""

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from oracle import fem_oracle as fo
from oracle import krylov_oracle as ko
from pygeon_b200.grid import structured_grid


def poisson(dim, n):
    sd = structured_grid(dim, n)
    sd.compute_geometry()
    A = fo.closed_form("lagrange1", "stiff", sd, rot2d=False).tocsr()
    keep = ~sd.tags["domain_boundary_nodes"]
    A = A[keep][:, keep].tocsr()
    b = (fo.closed_form("lagrange1", "mass", sd) @ np.ones(sd.num_nodes))[keep]
    return A, b


def darcy(dim, n):
    sd = structured_grid(dim, n)
    sd.compute_geometry()
    M = fo.closed_form("rt0", "mass", sd)
    B = fo.p0_mass(sd) @ fo.diff_matrix("rt0", sd)
    A = sps.block_array([[M, -B.T], [-B, None]]).tocsr()
    rng = np.random.default_rng(0)
    return A, rng.normal(size=A.shape[0])


@pytest.mark.parametrize("dim,n", [(2, 8), (3, 4)])
@pytest.mark.parametrize("jacobi", [False, True])
def test_cg_oracle_vs_scipy(dim, n, jacobi):
    A, b = poisson(dim, n)
    dinv = 1.0 / A.diagonal() if jacobi else None
    its = []
    M = None if dinv is None else spla.LinearOperator(A.shape, matvec=lambda v: dinv * v)
    xs, info_s = spla.cg(A, b, rtol=1e-10, atol=0.0, maxiter=500, M=M, callback=lambda xk: its.append(1))
    x, info, k, hist = ko.cg(A, b, rtol=1e-10, maxiter=500, dinv=dinv)
    assert info == info_s == 0 and k == len(its)
    assert np.allclose(x, xs, rtol=1e-12, atol=1e-14)
    assert hist[-1] < 1e-10 * np.linalg.norm(b)


@pytest.mark.parametrize("dim,n", [(2, 6), (3, 3)])
@pytest.mark.parametrize("precond", [False, True])
def test_minres_oracle_vs_scipy(dim, n, precond):
    A, b = darcy(dim, n)
    its = []
    dinv, M = None, None
    if precond:
        dinv = 1.0 / np.where(A.diagonal() > 0, A.diagonal(), 1.0 + np.arange(A.shape[0]) % 3)
        M = spla.LinearOperator(A.shape, matvec=lambda v: dinv * v)
    xs, info_s = spla.minres(A, b, rtol=1e-9, maxiter=4000, M=M, callback=lambda xk: its.append(1))
    x, info, k, hist = ko.minres(A, b, rtol=1e-9, maxiter=4000, dinv=dinv)
    assert info == info_s and k == len(its)
    assert np.allclose(x, xs, rtol=1e-10, atol=1e-12)
