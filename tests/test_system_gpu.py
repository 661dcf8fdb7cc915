"""System build in HBM: extraction kernels behind reduce_system / the
owned-row blocks, diagonal, block Jacobi, LinearSystem's device front end; fixes for zero
right-hand sides, iteration limits and non-symmetric tensors on the device-resident path."""

import time

import numpy as np
import pytest
import scipy.sparse as sps
import torch

import pygeon_b200 as pg
from oracle import fem_oracle as fo
from pgb_helpers import random_data

pytestmark = pytest.mark.gpu


def _host(A):
    return sps.csr_array((A.data.cpu().numpy(), A.indices.cpu().numpy(), A.indptr.cpu().numpy()), shape=A.shape)


@pytest.mark.parametrize("n,dens", [(1, 1.0), (257, 0.05), (5000, 0.004), (70001, 0.0002)])
@pytest.mark.parametrize("idx64", [False, True])
def test_extract_rows_and_columns(n, dens, idx64):
    """pgb_mask_to_map + pgb_csr_extract_* == scipy fancy indexing, entries in their original order;
    ragged sizes, empty rows, rows longer than a lane group, everything dropped."""
    from pygeon_b200 import engine
    from pygeon_b200.engine import DeviceCSR

    rng = np.random.default_rng(n)
    m = max(int(dens * n * n), 1)  # (sps.random draws without replacement from n^2 candidates: minutes at n = 7e4)
    A = sps.csr_array((rng.normal(size=m), (rng.integers(0, n, m), rng.integers(0, n, m))), shape=(n, n))
    A.sum_duplicates()
    if rng.random() > 0.5:
        A = sps.csr_array(A + sps.eye(n, format="csr"))
    if n > 300:  # one long row, some empty rows
        A = sps.csr_array(sps.vstack([sps.csr_array(np.ones((1, n))), A[1:]]))
        live = np.flatnonzero(rng.random(n) > 0.1)
        A = sps.csr_array(sps.csr_array((np.ones(live.size), (live, live)), shape=(n, n)) @ A)
    A.eliminate_zeros()
    A.sort_indices()
    Ad = DeviceCSR.from_scipy(A)
    if idx64:
        Ad = DeviceCSR(Ad.shape, Ad.indptr.to(torch.int64), Ad.indices, Ad.data)
    for frac in (0.7, 0.0, 1.0):
        keep = rng.random(n) < frac if 0 < frac < 1 else np.full(n, frac == 1.0)
        newid, rows = engine.mask_to_map(torch.from_numpy(keep).cuda())
        k = np.flatnonzero(keep)
        assert np.array_equal(rows.cpu().numpy(), k)
        ref_id = -np.ones(n, dtype=np.int64)
        ref_id[k] = np.arange(k.size)
        assert np.array_equal(newid.cpu().numpy(), ref_id)
        sub = _host(Ad.extract(rows, newid, k.size))
        ref = sps.csr_array(A[k][:, k])
        ref.sort_indices()
        assert np.array_equal(sub.indptr, ref.indptr) and np.array_equal(sub.indices, ref.indices)
        assert np.array_equal(sub.data, ref.data)
    # non-monotone column map (owned | ghost renumbering): order of the entries inside a row is kept
    perm = rng.permutation(n).astype(np.int32)
    sub = Ad.extract(None, torch.from_numpy(perm).cuda(), n)
    assert np.array_equal(sub.indices.cpu().numpy(), perm[A.indices])
    assert np.array_equal(sub.data.cpu().numpy(), A.data)


def test_diagonal_and_block_jacobi():
    sd = pg.structured_grid(3, 5, perturb=0.1)
    sd.compute_geometry()
    data = random_data(sd.num_cells, seed=2)
    Ad = pg.darcy_saddle_device(sd, data, "key")
    A = Ad.to_scipy("csr")
    assert np.array_equal(Ad.diagonal().cpu().numpy(), A.diagonal())
    nf = sd.num_faces
    dinv = pg.darcy_block_jacobi(Ad, nf).cpu().numpy()
    du = A.diagonal()[:nf]
    B = A[nf:, :nf]
    dp = np.asarray((B.multiply(B) @ (1.0 / du))).ravel()
    assert np.allclose(dinv, 1.0 / np.concatenate([du, dp]), rtol=1e-13)


def test_cg_zero_rhs_returns_at_once():
    """b == 0 with the default maxiter (10 n) must not iterate (scipy returns b)."""
    sd = pg.structured_grid(3, 24)
    sd.compute_geometry()
    A = pg.Lagrange1().assemble_mass_matrix_device(sd)
    t = time.perf_counter()
    x, info = pg.cg_device(A, torch.zeros(sd.num_nodes, dtype=torch.float64, device="cuda"))
    torch.cuda.synchronize()
    assert time.perf_counter() - t < 5.0
    assert info.info == 0 and info.iterations == 0 and not bool(x.any())
    x, info = pg.minres_device(A, torch.zeros(sd.num_nodes, dtype=torch.float64, device="cuda"))
    assert info.info == 0 and info.iterations == 0 and not bool(x.any())


def test_maxiter_range_and_history_cap():
    from pygeon_b200 import solvers

    assert solvers._maxiter(None, 10 * 3 * 10**8) == 2**31 - 1
    with pytest.raises(ValueError):
        solvers._maxiter(2**31, 5)
    with pytest.raises(ValueError):
        solvers._maxiter(-1, 5)
    # a huge iteration limit allocates a bounded history (no 8 * maxiter bytes on host or device)
    sd = pg.structured_grid(2, 8)
    sd.compute_geometry()
    A = pg.Lagrange1().assemble_mass_matrix_device(sd)
    b = torch.ones(sd.num_nodes, dtype=torch.float64, device="cuda")
    free0 = torch.cuda.mem_get_info()[0]
    x, info = pg.cg_device(A, b, rtol=1e-12, maxiter=2**31 - 1)
    assert info.info == 0 and info.residuals.size == info.iterations + 1
    assert free0 - torch.cuda.mem_get_info()[0] < 64 << 20
    x, info = pg.minres_device(A, b, rtol=1e-12, maxiter=2**31 - 1)
    assert info.info == 0 and info.residuals.size == info.iterations


@pytest.mark.parametrize("dim,n", [(2, 9), (3, 4)])
def test_nonsymmetric_tensor_on_the_device_path(dim, n):
    """the device-resident matrix is consumed as CSR (SpMV, reduce_system, Krylov).  With a
    non-symmetric K it must be the reference's matrix, not its transpose; the CSC host copy too."""
    sd = pg.structured_grid(dim, n, perturb=0.2)
    sd.compute_geometry()
    data = random_data(sd.num_cells, seed=4, sym=False)
    x = np.random.default_rng(1).normal(size=sd.num_faces)
    A_ref = fo.closed_form("rt0", "mass", sd, data, "key")  # the reference's matrix (pinned by the goldens)
    assert abs(A_ref - A_ref.T).max() > 1e-3 * abs(A_ref).max()
    Ad = pg.RT0("key").assemble_mass_matrix_device(sd, data)
    assert not Ad.symmetric
    y = Ad.matvec(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(y - A_ref @ x).max() <= 1e-12 * np.abs(A_ref @ x).max()
    for fmt in ("csc", "csr"):
        H = Ad.to_scipy(fmt)
        assert abs(H - A_ref).max() <= 1e-12 * abs(A_ref).max()
    H = pg.RT0("key").assemble_mass_matrix(sd, data)
    assert abs(H - A_ref).max() <= 1e-12 * abs(A_ref).max()
    S = pg.darcy_saddle_device(sd, data, "key")
    M = S.to_scipy("csr")[: sd.num_faces, : sd.num_faces]
    assert abs(M - A_ref).max() <= 1e-12 * abs(A_ref).max()


def test_linear_system_device_front_end():
    """pg.LinearSystem.solve(pg.cg_solver(...)) eliminates and solves in HBM; same answer as the
    host route with the same solver, for a scipy matrix and for a DeviceCSR, one and several rhs."""
    sd = pg.structured_grid(3, 6, perturb=0.1)
    sd.compute_geometry()
    A = pg.Lagrange1().assemble_stiff_matrix(sd)
    Ad = pg.Lagrange1().assemble_stiff_matrix_device(sd)
    rng = np.random.default_rng(5)
    bd = sd.tags["domain_boundary_nodes"]
    vals = rng.normal(size=sd.num_nodes)
    for b in (rng.normal(size=sd.num_nodes), rng.normal(size=(sd.num_nodes, 2))):
        ref = pg.LinearSystem(A, b)
        ref.flag_ess_bc(bd, vals)
        x_ref = ref.solve()  # spsolve
        for mat in (A, Ad):
            ls = pg.LinearSystem(mat, b)
            ls.flag_ess_bc(bd, vals)
            solver = pg.cg_solver(rtol=1e-13, maxiter=3000)
            x = ls.solve(solver)
            assert x.shape == x_ref.shape and np.abs(x - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
            info = solver.info if b.ndim == 1 else solver.info[0]
            assert info.info == 0 and info.iterations > 5
        # plain callable use of the same solver object (scipy-style)
        A0, b0, _ = ref.reduce_system()
        x0 = pg.cg_solver(rtol=1e-13)(A0, b0 if b.ndim == 1 else b0[:, 0])
        assert np.abs(A0 @ x0 - (b0 if b.ndim == 1 else b0[:, 0])).max() < 1e-9


def test_mass_dispatchers_on_mdg():
    """pg.cell_mass / face_mass / ridge_mass / peak_mass, lumped variants and *_stiff on a
    one-subdomain mixed-dimensional grid equal the discretization calls
    (numerics/innerproducts.py:13-160, stiffness.py:80-109; tests/numerics/test_innerproducts.py:16-52)."""
    sd = pg.structured_grid(3, 3, perturb=0.1)
    sd.compute_geometry()
    mdg = pg.as_mdg(sd)
    assert np.allclose(pg.cell_mass(mdg) @ sd.cell_volumes, 1)
    pairs = [(pg.cell_mass, pg.lumped_cell_mass, pg.PwConstants, mdg.num_subdomain_cells()),
             (pg.face_mass, pg.lumped_face_mass, pg.RT0, mdg.num_subdomain_faces()),
             (pg.ridge_mass, pg.lumped_ridge_mass, pg.Nedelec0, mdg.num_subdomain_ridges()),
             (pg.peak_mass, pg.lumped_peak_mass, pg.Lagrange1, mdg.num_subdomain_peaks())]
    for mass, lumped, cls, n in pairs:
        M, L = mass(mdg), lumped(mdg)
        assert M.shape == L.shape == (n, n)
        assert abs(M - cls().assemble_mass_matrix(sd)).max() == 0
        assert abs(L - cls().assemble_lumped_matrix(sd)).max() <= 1e-15
        assert abs(M - M.T).max() <= 1e-15 * abs(M).max()
    assert abs(pg.face_mass(mdg, pg.BDM1()) - pg.BDM1().assemble_mass_matrix(sd)).max() == 0
    assert pg.cell_stiff(mdg).nnz == 0
    for stiff, cls in ((pg.face_stiff, pg.RT0), (pg.ridge_stiff, pg.Nedelec0), (pg.peak_stiff, pg.Lagrange1)):
        S, ref = stiff(mdg), cls().assemble_stiff_matrix(sd)
        assert abs(S - ref).max() <= 1e-12 * abs(ref).max()
