"""Parity of the CUDA SpMV / CG / MINRES with scipy (the oracle for the solve, see
oracle/krylov_oracle.py) and the reference's patch tests (exact solutions)."""

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla
import torch

import pygeon_b200 as pg
from oracle import krylov_oracle as ko

pytestmark = pytest.mark.gpu


def poisson_system(dim, n):
    sd = pg.structured_grid(dim, n)
    sd.compute_geometry()
    A = pg.Lagrange1().assemble_grad_grad_matrix(sd)
    M = pg.Lagrange1().assemble_mass_matrix(sd)
    ls = pg.LinearSystem(A, M @ np.ones(sd.num_nodes))
    bd = sd.tags["domain_boundary_nodes"]
    ls.flag_ess_bc(bd, np.zeros(sd.num_nodes))
    return sd, ls


def darcy_system(dim, n):
    sd = pg.structured_grid(dim, n)
    sd.compute_geometry()
    M = pg.RT0().assemble_mass_matrix(sd)
    B = pg.PwConstants().assemble_mass_matrix(sd) @ pg.RT0().assemble_diff_matrix(sd)
    A = sps.block_array([[M, -B.T], [-B, None]]).tocsc()
    return sd, A, M, B


@pytest.mark.parametrize("dim,n", [(2, 5), (2, 100), (3, 17)])
def test_spmv(dim, n):
    from pygeon_b200 import _lib
    from pygeon_b200.engine import DeviceCSR

    sd = pg.structured_grid(dim, n, perturb=0.1)
    sd.compute_geometry()
    for disc in (pg.Lagrange1(), pg.RT0(), pg.Nedelec0()):
        A = sps.csr_array(disc.assemble_mass_matrix(sd))
        Ad = DeviceCSR.from_scipy(A)
        x = np.random.default_rng(0).normal(size=A.shape[1])
        xd = torch.from_numpy(x).cuda()
        for tpr in (0, 2, 4, 8, 16, 32):
            yd = torch.empty_like(xd)
            _lib.check(_lib.lib.pgb_spmv(A.shape[0], Ad.indptr.data_ptr(), 0, Ad.indices.data_ptr(),
                                         Ad.data.data_ptr(), xd.data_ptr(), yd.data_ptr(), tpr,
                                         torch.cuda.current_stream().cuda_stream))
            y = yd.cpu().numpy()
            assert np.abs(y - A @ x).max() <= 1e-13 * np.abs(A @ x).max()
    # int64 indptr path
    Ad64 = DeviceCSR(Ad.shape, Ad.indptr.to(torch.int64), Ad.indices, Ad.data)
    yd = torch.empty_like(xd)
    _lib.check(_lib.lib.pgb_spmv(A.shape[0], Ad64.indptr.data_ptr(), 1, Ad64.indices.data_ptr(),
                                 Ad64.data.data_ptr(), xd.data_ptr(), yd.data_ptr(), 0,
                                 torch.cuda.current_stream().cuda_stream))
    assert np.abs(yd.cpu().numpy() - A @ x).max() <= 1e-13 * np.abs(A @ x).max()


@pytest.mark.parametrize("dim,n", [(2, 24), (3, 8), (3, 16)])
@pytest.mark.parametrize("jacobi", [False, True])
def test_cg_matches_scipy(dim, n, jacobi):
    """Same x0 = 0, same stopping rule: identical iteration count, residual history within 1e-10
    relative, final iterate within 1e-10 (north_star bar)."""
    sd, ls = poisson_system(dim, n)
    A0, b0, _ = ls.reduce_system()
    A0 = sps.csr_array(A0)
    dinv = 1.0 / A0.diagonal() if jacobi else None
    x_o, info_o, k_o, hist_o = ko.cg(A0, b0, rtol=1e-10, maxiter=2000, dinv=dinv)
    x_g, info_g = pg.cg(A0, b0, rtol=1e-10, maxiter=2000, M=dinv)
    si = pg.cg.last_info
    assert info_g == info_o == 0
    assert si.iterations == k_o
    assert np.abs(si.residuals - hist_o).max() <= 1e-10 * hist_o[0]
    assert np.abs(x_g - x_o).max() <= 1e-10 * np.abs(x_o).max()
    # and against scipy itself
    Mop = None if dinv is None else spla.LinearOperator(A0.shape, matvec=lambda v: dinv * v)
    x_s, info_s = spla.cg(A0, b0, rtol=1e-10, atol=0.0, maxiter=2000, M=Mop)
    assert info_s == 0 and np.abs(x_g - x_s).max() <= 1e-10 * np.abs(x_s).max()


def test_cg_maxiter_and_zero_rhs_and_x0():
    sd, ls = poisson_system(2, 16)
    A0, b0, _ = ls.reduce_system()
    x, info = pg.cg(A0, b0, rtol=1e-14, maxiter=5)
    xo, io, ko_, hist = ko.cg(sps.csr_array(A0), b0, rtol=1e-14, maxiter=5)
    assert info == io == 5 and pg.cg.last_info.iterations == 5
    assert np.abs(x - xo).max() <= 1e-12 * np.abs(xo).max()
    x, info = pg.cg(A0, np.zeros_like(b0))
    assert info == 0 and not x.any()
    x0 = np.random.default_rng(0).normal(size=b0.size)
    x, info = pg.cg(A0, b0, x0=x0, rtol=1e-11, maxiter=500)
    xo, io, _, _ = ko.cg(sps.csr_array(A0), b0, x0=x0, rtol=1e-11, maxiter=500)
    assert info == io == 0 and np.abs(x - xo).max() <= 1e-9 * np.abs(xo).max()


@pytest.mark.parametrize("dim,n", [(2, 8), (3, 4)])
def test_minres_matches_scipy(dim, n):
    sd, A, M, B = darcy_system(dim, n)
    rng = np.random.default_rng(1)
    b = rng.normal(size=A.shape[0])
    x_o, info_o, k_o, hist_o = ko.minres(sps.csr_array(A), b, rtol=1e-9, maxiter=20000)
    x_g, info_g = pg.minres(A, b, rtol=1e-9, maxiter=20000)
    si = pg.minres.last_info
    assert info_g == info_o
    # long recurrences amplify rounding differences of the dot products: allow a few iterations
    assert abs(si.iterations - k_o) <= max(3, k_o // 50)
    m = min(si.iterations, k_o, 50)
    assert np.abs(si.residuals[:m] - hist_o[:m]).max() <= 1e-9 * hist_o[0]
    r = b - A @ x_g
    ro = b - A @ x_o
    assert np.linalg.norm(r) <= 10 * max(np.linalg.norm(ro), 1e-9 * np.linalg.norm(b))


def test_minres_spd_and_shift():
    sd, ls = poisson_system(2, 12)
    A0, b0, _ = ls.reduce_system()
    A0 = sps.csr_array(A0)
    for shift in (0.0, -0.5):
        x_s, info_s = spla.minres(A0, b0, rtol=1e-11, shift=shift, maxiter=3000)
        x_g, info_g = pg.minres(A0, b0, rtol=1e-11, shift=shift, maxiter=3000)
        assert info_g == info_s == 0
        assert np.abs(x_g - x_s).max() <= 1e-8 * np.abs(x_s).max()


def test_patch_poisson_quadratic_exact():
    """Reference patch test (tests/patch_tests/test_laplacian.py:19-55): -lap u = f with a solution
    in the discrete space is reproduced; here a linear u with Dirichlet data on the boundary."""
    sd = pg.structured_grid(3, 6, perturb=0.15)
    sd.compute_geometry()
    A = pg.Lagrange1().assemble_stiff_matrix(sd)
    u_ex = 1.0 + 2 * sd.nodes[0] - 3 * sd.nodes[1] + 0.5 * sd.nodes[2]
    ls = pg.LinearSystem(A, np.zeros(sd.num_nodes))
    ls.flag_ess_bc(sd.tags["domain_boundary_nodes"], u_ex)
    u = ls.solve(pg.cg_solver(rtol=1e-13, maxiter=2000))
    assert np.abs(u - u_ex).max() < 1e-9
    u_direct = ls.solve()  # scipy spsolve, the reference's default
    assert np.abs(u - u_direct).max() < 1e-9


def test_patch_mixed_darcy_linear_pressure():
    """Reference patch test (tests/patch_tests/test_laplacian_mixed.py:27-70): RT0-P0 reproduces a
    linear pressure; solved with GPU MINRES on the symmetrised saddle system."""
    sd = pg.structured_grid(2, 6)
    sd.compute_geometry()
    M = pg.RT0().assemble_mass_matrix(sd)
    P0 = pg.PwConstants().assemble_mass_matrix(sd)
    div = pg.RT0().assemble_diff_matrix(sd)
    B = P0 @ div
    A = sps.block_array([[M, -B.T], [-B, None]]).tocsc()
    # natural bc for p = x on all boundary faces: -(u.n, p)_bd  (hdiv.py:214-241)
    bf = sd.tags["domain_boundary_faces"]
    sign = np.asarray(sd.cell_faces.sum(axis=1)).ravel()
    rhs_u = np.zeros(sd.num_faces)
    rhs_u[bf] = -sd.face_centers[0, bf] * sign[bf]
    b = np.concatenate([rhs_u, np.zeros(sd.num_cells)])
    x, info = pg.minres(A, b, rtol=1e-13, maxiter=20000)
    assert info == 0
    u, p = x[: sd.num_faces], x[sd.num_faces:]
    # P0 dofs are cell integrals
    assert np.abs(p / sd.cell_volumes - sd.cell_centers[0]).max() < 1e-7
    # flux u = -grad p = (-1, 0): face dof = -n_x
    assert np.abs(u + sd.face_normals[0]).max() < 1e-7


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("space", ["RT0", "BDM1"])
def test_patch_linear_distribution_reference_script(dim, space):
    """The reference's patch test, statement by statement (tests/patch_tests/
    test_laplacian_mixed.py:27-70): saddle system from assemble_mass_matrix / assemble_diff_matrix,
    natural boundary data from assemble_nat_bc, exact fields from interpolate."""
    sd = pg.structured_grid(dim, 4 if dim == 2 else 3)
    sd.compute_geometry()
    discr_q, discr_p = getattr(pg, space)("flow"), pg.PwConstants("flow")

    def q_func(_):
        return np.array([-1, 2, 1 if dim == 3 else 0])

    def p_func(x):
        return -x @ q_func(x)

    face_mass = discr_q.assemble_mass_matrix(sd)
    cell_mass = discr_p.assemble_mass_matrix(sd, None)
    div = cell_mass @ discr_q.assemble_diff_matrix(sd)
    spp = sps.block_array([[face_mass, -div.T], [div, None]]).tocsc()
    b_faces = sd.tags["domain_boundary_faces"]
    bc_val = -discr_q.assemble_nat_bc(sd, p_func, b_faces)
    rhs = np.zeros(spp.shape[0])
    rhs[: bc_val.size] += bc_val
    x = pg.LinearSystem(spp, rhs).solve()  # scipy spsolve, the reference's default
    q, p = x[: bc_val.size], x[-discr_p.ndof(sd):]
    assert np.allclose(p, discr_p.interpolate(sd, p_func))
    assert np.allclose(q, discr_q.interpolate(sd, q_func))
    # the same system, symmetrised, through GPU MINRES
    sym = sps.block_array([[face_mass, -div.T], [-div, None]]).tocsc()
    xs, info = pg.minres(sym, rhs, rtol=1e-13, maxiter=50000)
    assert info == 0 and np.abs(xs - x).max() < 1e-6 * np.abs(x).max()


def test_linear_system_multi_rhs():
    """b with several columns (tests/numerics/test_linear_system.py:9-24)."""
    sd, ls = poisson_system(2, 10)
    n = sd.num_nodes
    b2 = np.stack([ls.b, 2 * ls.b], axis=1)
    ls2 = pg.LinearSystem(ls.A, b2)
    ls2.flag_ess_bc(sd.tags["domain_boundary_nodes"], np.zeros(n))
    x = ls2.solve(pg.cg_solver(rtol=1e-12))
    assert x.shape == (n, 2) and np.allclose(x[:, 1], 2 * x[:, 0], rtol=1e-9, atol=1e-12)


def test_device_linear_system_matches_host():
    """Assemble -> essential BC elimination -> CG entirely on the device equals the host
    LinearSystem route (numerics/linear_system.py:64-94)."""
    sd = pg.structured_grid(3, 7, perturb=0.1)
    sd.compute_geometry()
    A = pg.Lagrange1().assemble_stiff_matrix(sd)
    Ad = pg.Lagrange1().assemble_stiff_matrix_device(sd)
    rng = np.random.default_rng(3)
    b = rng.normal(size=sd.num_nodes)
    bd = sd.tags["domain_boundary_nodes"]
    vals = rng.normal(size=sd.num_nodes)
    ls = pg.LinearSystem(A, b)
    ls.flag_ess_bc(bd, vals)
    x_host = ls.solve()
    dls = pg.DeviceLinearSystem(Ad, b)
    dls.flag_ess_bc(bd, vals)
    A0, b0, keep = dls.reduce_system()
    A0h, b0h, _ = ls.reduce_system()
    A0s = sps.csr_array((A0.data.cpu().numpy(), A0.indices.cpu().numpy(), A0.indptr.cpu().numpy()), shape=A0.shape)
    assert abs(A0s - sps.csr_array(A0h)).max() <= 1e-14 * abs(A0h).max()
    assert np.abs(b0.cpu().numpy() - b0h).max() <= 1e-13 * np.abs(b0h).max()
    x, info = dls.solve("cg", rtol=1e-13, maxiter=2000)
    assert info.info == 0 and np.abs(x.cpu().numpy() - x_host).max() <= 1e-9 * np.abs(x_host).max()


@pytest.mark.parametrize("dim,n", [(2, 12), (3, 4)])
def test_preconditioned_minres_darcy(dim, n):
    """Block-Jacobi preconditioned MINRES on the one-pass saddle system: same recurrence as scipy's
    minres with M (oracle), and far fewer iterations than without."""
    sd = pg.structured_grid(dim, n, perturb=0.1)
    sd.compute_geometry()
    Ad = pg.darcy_saddle_device(sd)
    dinv = pg.darcy_block_jacobi(Ad, sd.num_faces)
    A = Ad.to_scipy("csr")
    b = np.random.default_rng(2).normal(size=A.shape[0])
    x_o, info_o, k_o, hist_o = ko.minres(A, b, rtol=1e-10, maxiter=20000, dinv=dinv.cpu().numpy())
    x_g, info_g = pg.minres(Ad, b, rtol=1e-10, maxiter=20000, M=dinv)
    si = pg.minres.last_info
    assert info_g == info_o == 0
    assert abs(si.iterations - k_o) <= max(3, k_o // 50)
    m = min(si.iterations, k_o, 40)
    assert np.abs(si.residuals[:m] - hist_o[:m]).max() <= 1e-9 * hist_o[0]
    _, _ = pg.minres(Ad, b, rtol=1e-10, maxiter=20000)
    assert pg.minres.last_info.iterations > 2 * si.iterations
    r = np.linalg.norm(b - A @ x_g)
    r_o = np.linalg.norm(b - A @ x_o)
    assert r <= 10 * r_o  # MINRES stops on |r| <= rtol |A| |x|: compare with the oracle, not with |b|
