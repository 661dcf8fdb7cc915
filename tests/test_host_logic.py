"""Host-side logic that needs no GPU: coefficient convention, LinearSystem, grid container."""

import numpy as np
import pytest
import scipy.sparse as sps

import pygeon_b200 as pg
from pygeon_b200.linear_system import LinearSystem, create_restriction
from pygeon_b200.params import coefficient as get_cell_data


def test_get_cell_data_defaults_and_broadcast():
    sd = pg.structured_grid(2, 2)
    nc = sd.num_cells
    assert get_cell_data(sd, None, "k", pg.WEIGHT) is None
    assert get_cell_data(sd, {}, "k", pg.SECOND_ORDER_TENSOR) is None
    data = pg.initialize_data({}, "k", {pg.WEIGHT: 2.0})
    w = get_cell_data(sd, data, "k", pg.WEIGHT)
    assert w.shape == (nc,) and np.all(w == 2.0)
    assert get_cell_data(sd, data, "other", pg.WEIGHT) is None  # unknown keyword -> default (data.py:94-100)
    sot = pg.SecondOrderTensor(np.arange(1.0, nc + 1), kxy=np.ones(nc))
    data = pg.initialize_data({}, "k", {pg.SECOND_ORDER_TENSOR: sot})
    K = get_cell_data(sd, data, "k", pg.SECOND_ORDER_TENSOR)
    assert K.shape == (3, 3, nc) and np.all(K[0, 1] == 1) and np.all(K[2, 2] == K[0, 0])
    with pytest.raises(ValueError):
        get_cell_data(sd, pg.initialize_data({}, "k", {pg.WEIGHT: np.ones(nc + 1)}), "k", pg.WEIGHT)


def test_linear_system_matches_direct_elimination():
    """R A R^T, R(b - A ess), ess + R^T x0 (numerics/linear_system.py:64-94), single and multi RHS."""
    rng = np.random.default_rng(0)
    n = 12
    A = sps.random(n, n, 0.4, random_state=1) + sps.eye(n) * 5
    A = sps.csc_array(A + A.T)
    b = rng.normal(size=n)
    ess = np.zeros(n, dtype=bool)
    ess[[0, 5, 11]] = True
    vals = rng.normal(size=n)
    ls = LinearSystem(A, b)
    ls.flag_ess_bc(ess, vals)
    x = ls.solve()
    assert np.allclose(x[ess], vals[ess])
    assert np.allclose((A @ x)[~ess], b[~ess])
    R = create_restriction(~ess)
    assert R.shape == (n - 3, n) and np.array_equal((R @ np.arange(n)), np.arange(n)[~ess])
    b2 = np.stack([b, -b], axis=1)
    ls2 = LinearSystem(A, b2)
    ls2.flag_ess_bc(ess, np.zeros(n))
    x2 = ls2.solve()
    assert x2.shape == (n, 2) and np.allclose(x2[:, 0], -x2[:, 1])
    ls2.reset_bc()
    assert ls2.is_dof.all()


def test_structured_grid_counts():
    """Entity counts of the structured simplicial grids (tests/grids/test_grid_ridges.py:44-73 of
    the reference: N = 1 tet cube has 19 edges) and SURVEY section 8 formulas."""
    sd = pg.structured_grid(3, 1)
    sd.compute_geometry()
    assert (sd.num_cells, sd.num_faces, sd.num_ridges, sd.num_nodes) == (6, 18, 19, 8)
    assert sorted(set(np.round(sd.edge_lengths**2).astype(int))) == [1, 2, 3]
    for n in (2, 5):
        sd = pg.structured_grid(3, n)
        sd.compute_connectivity()
        assert sd.num_cells == 6 * n**3 and sd.num_faces == 12 * n**3 + 6 * n**2
        assert sd.num_ridges == 7 * n**3 + 9 * n**2 + 3 * n and sd.num_nodes == (n + 1) ** 3
        s2 = pg.structured_grid(2, n)
        s2.compute_geometry()
        assert s2.num_cells == 2 * n**2 and s2.num_faces == 3 * n**2 + 2 * n and s2.num_edges == s2.num_faces


def test_non_simplicial_grid_raises():
    sd = pg.structured_grid(2, 2)
    cf = sd.cell_faces.tolil()
    bad = pg.SimplexGrid(2, sd.nodes, sd.face_nodes, sps.csc_array(sps.hstack([cf[:, :1] + cf[:, 1:2], cf[:, 2:]])))
    with pytest.raises(NotImplementedError):
        bad.compute_geometry()


def test_api_surface_matches_reference_names():
    for name in ("Lagrange1", "RT0", "BDM1", "Nedelec0", "PwConstants", "LinearSystem", "UNITARY_DATA",
                 "SECOND_ORDER_TENSOR", "WEIGHT", "AMBIENT_DIM", "SCALAR", "VECTOR", "MATRIX"):
        assert hasattr(pg, name), name
    d = pg.RT0()
    assert d.keyword == pg.UNITARY_DATA and repr(d) == "Discretization of type RT0 with keyword unitary_data"
    assert pg.RT0.poly_order == 1 and pg.RT0.tensor_order == pg.VECTOR and pg.Lagrange1.tensor_order == pg.SCALAR
    assert pg.Lagrange1().get_range_discr_class(3) is pg.Nedelec0 and pg.Lagrange1().get_range_discr_class(2) is pg.RT0
    assert pg.Nedelec0().get_range_discr_class(3) is pg.RT0 and pg.RT0().get_range_discr_class(3) is pg.PwConstants
    with pytest.raises(NotImplementedError):
        pg.Lagrange1().get_range_discr_class(0)
