"""Operator layer above the discretizations (no GPU): exterior derivatives, block helpers, tip
restrictions, the mixed-dimensional container and LinearSystem's host path, against the unmodified
reference where it is importable and against the reference's own test expectations otherwise
(tests/numerics/test_differentials.py:13-52, test_linear_system.py:9-24)."""

import numpy as np
import pytest
import scipy.sparse as sps

import pygeon_b200 as pg
from oracle import ref_loader


def _cochain(grid):
    grad, curl, div = pg.grad(grid), pg.curl(grid), pg.div(grid)
    assert (curl @ grad).nnz == 0 and (div @ curl).nnz == 0


@pytest.mark.parametrize("dim,n", [(2, 3), (3, 2)])
def test_cochain_and_shapes(dim, n):
    sd = pg.structured_grid(dim, n)
    sd.compute_geometry()
    _cochain(sd)
    ext = pg.numerics.differentials.exterior_derivative
    assert ext(sd, 4).shape == (sd.num_peaks, 0) and ext(sd, 0).shape == (0, sd.num_cells)
    with pytest.warns():
        ext(sd, -1)
    with pytest.raises(TypeError):
        pg.grad([])
    mdg = pg.as_mdg(sd)
    assert mdg.num_subdomains() == 1 and mdg.dim_max() == dim and mdg.num_subdomain_cells() == sd.num_cells
    sd.tags.setdefault("tip_ridges", np.zeros(sd.num_ridges, dtype=bool))
    sd.tags.setdefault("tip_peaks", np.zeros(sd.num_peaks, dtype=bool))
    _cochain(mdg)
    assert abs(pg.div(mdg) - pg.div(sd)).nnz == 0
    blk = pg.div(mdg, as_bmat=True)
    assert blk.shape == (1, 1) and abs(blk[0, 0] - pg.div(sd)).nnz == 0
    assert pg.remove_tip_dofs(mdg, 1).shape == (sd.num_faces, sd.num_faces)


@pytest.mark.skipif(not ref_loader.available(), reason="reference source not present")
@pytest.mark.parametrize("dim,n", [(2, 3), (3, 2)])
def test_differentials_match_reference(dim, n):
    rpg, _ = ref_loader.load()
    sd = pg.structured_grid(dim, n, perturb=0.1)
    sd.compute_geometry()
    g = ref_loader.ref_grid(sd)
    for name in ("div", "curl", "grad"):
        ours, ref = getattr(pg, name)(sd), getattr(rpg, name)(g)
        assert ours.shape == ref.shape and abs(sps.csr_array(ours) - sps.csr_array(ref)).nnz == 0


def test_bmat_helpers():
    A = np.empty((2, 2), dtype=object)
    A[0, 0], A[1, 1] = sps.eye_array(2, format="csc") * 2, sps.eye_array(3, format="csc") * 3
    pg.bmat.replace_nones_with_zeros(A)
    assert A[0, 1].shape == (2, 3) and A[1, 0].shape == (3, 2)
    rows, cols = pg.bmat.find_row_col_lengths(A)
    assert list(rows) == [2, 3] and list(cols) == [2, 3]
    P = pg.bmat.multiply(A, pg.bmat.transpose(A))
    assert np.allclose(sps.block_array(P).toarray(), np.diag([4, 4, 9, 9, 9]))


def test_default_discr():
    sd = pg.structured_grid(3, 1)
    dd = pg.numerics.innerproducts.default_discr
    assert [type(dd(sd, k)).__name__ for k in range(4)] == ["PwConstants", "RT0", "Nedelec0", "Lagrange1"]
    assert type(dd(pg.structured_grid(2, 1), 2)).__name__ == "Lagrange1"
    assert dd(sd, 1, keyword="flow").keyword == "flow"
    with pytest.raises(ValueError):
        dd(sd, -1)


def test_linear_system_reference_case():
    """tests/numerics/test_linear_system.py:9-24 of the reference."""
    A = sps.eye(4, 4).tocsc()
    b = np.tile(np.arange(4), (3, 1)).T
    LS = pg.LinearSystem(A, b)
    assert np.allclose(LS.solve(), b)
    ess_dofs = np.array([True, False, False, True])
    LS.flag_ess_bc(ess_dofs, np.ones_like(ess_dofs))
    assert np.allclose(LS.solve(), np.tile([1, 1, 2, 1], (3, 1)).T)


@pytest.mark.skipif(not ref_loader.available(), reason="reference source not present")
def test_linear_system_matches_reference():
    rpg, _ = ref_loader.load()
    rng = np.random.default_rng(3)
    n = 30
    A = sps.random(n, n, 0.2, random_state=2) + sps.eye(n) * 4
    A = sps.csc_array(A + A.T)
    ess = rng.random(n) < 0.25
    vals = rng.normal(size=n)
    for b in (rng.normal(size=n), rng.normal(size=(n, 3))):
        ours, ref = pg.LinearSystem(A, b), rpg.LinearSystem(A, b)
        for ls in (ours, ref):
            ls.flag_ess_bc(ess, vals)
            ls.flag_ess_bc(ess, 0.5 * vals)  # accumulates like the reference
        A0, b0, R0 = ours.reduce_system()
        A0r, b0r, R0r = ref.reduce_system()
        assert abs(sps.csr_array(A0) - sps.csr_array(A0r)).max() < 1e-14
        assert np.allclose(b0, np.asarray(b0r), atol=1e-14) and abs(R0 - R0r).nnz == 0
        assert np.allclose(ours.solve(), np.asarray(ref.solve()), atol=1e-12)


def test_get_cell_data_reference_semantics():
    """params/data.py:61-118: scalars broadcast, matrices wrapped, vectors tiled."""
    sd = pg.structured_grid(2, 2)
    nc = sd.num_cells
    assert np.array_equal(pg.get_cell_data(sd, None, "k", pg.WEIGHT), np.ones(nc))
    K = pg.get_cell_data(sd, {}, "k", pg.SECOND_ORDER_TENSOR, pg.MATRIX)
    assert K.values.shape == (3, 3, nc) and np.array_equal(K.values[:, :, 0], np.eye(3))
    data = pg.initialize_data({}, "k", {pg.VECTOR_FIELD: np.array([1.0, 2.0, 3.0])})
    v = pg.get_cell_data(sd, data, "k", pg.VECTOR_FIELD, pg.VECTOR)
    assert v.shape == (3, nc) and np.all(v[1] == 2.0)
    assert np.array_equal(pg.get_cell_data(sd, data, "k", pg.NORMAL_DIFFUSIVITY), np.ones(nc))


def test_tilted_grid_geometry_and_rotation():
    """A 2D grid moved rigidly into a generic plane: orthonormal rotation_matrix mapping the nodes
    back to a plane, in-plane face normals, unchanged measures (grids/grid.py:355-370)."""
    from oracle.gen_golden import tilt

    flat = pg.structured_grid(2, 3, perturb=0.2)
    flat.compute_geometry()
    sd = tilt(pg.structured_grid(2, 3, perturb=0.2))
    sd.compute_geometry()
    R = sd.rotation_matrix
    assert R.shape == (2, 3) and np.allclose(R @ R.T, np.eye(2), atol=1e-15)
    n_plane = np.cross(R[0], R[1])
    assert np.abs(n_plane @ (sd.nodes - sd.nodes[:, :1])).max() < 1e-14
    assert np.allclose(sd.cell_volumes, flat.cell_volumes, rtol=1e-13)
    assert np.allclose(sd.face_areas, flat.face_areas, rtol=1e-13)
    assert np.abs(n_plane @ sd.face_normals).max() < 1e-14
