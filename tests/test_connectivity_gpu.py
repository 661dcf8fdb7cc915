"""Device connectivity / geometry: ``pgb_ridges3d``, ``pgb_grid_geometry``,
``pgb_edge_geometry`` against the numpy restatement in ``SimplexGrid`` (itself pinned against the
reference's ``pg.Grid.compute_ridges`` / ``compute_edge_properties`` in
``tests/test_oracle_vs_reference.py::test_grid_contract_matches_reference``)."""

import numpy as np
import pytest
import scipy.sparse as sps

import pygeon_b200 as pg
from pgb_helpers import delaunay_grid
from pygeon_b200.grid import SimplexGrid

pytestmark = pytest.mark.gpu

GEOM = ("face_normals", "face_areas", "face_centers", "cell_centers", "cell_volumes", "edge_tangents", "edge_lengths")


def _raw_copy(sd):
    """The same grid as a plain SimplexGrid (no structured side tables, nothing computed)."""
    return SimplexGrid(sd.dim, sd.nodes.copy(), sd.face_nodes.copy(), sd.cell_faces.copy(), "copy")


def _grids():
    yield "delaunay3d", delaunay_grid(3, 400, seed=3)
    yield "delaunay2d", delaunay_grid(2, 300, seed=4)
    yield "kuhn3d", _raw_copy(pg.structured_grid(3, 5, perturb=0.2))
    yield "kuhn2d", _raw_copy(pg.structured_grid(2, 7, perturb=0.2))
    yield "reference_tet", pg.grid.reference_element(3)


@pytest.mark.parametrize("name", ["delaunay3d", "delaunay2d", "kuhn3d", "kuhn2d", "reference_tet"])
def test_device_geometry_matches_host(name):
    sd = dict(_grids())[name]
    host, dev = _raw_copy(sd), _raw_copy(sd)
    host.compute_geometry()
    dev.compute_geometry(device="cuda:0")
    for a in ("num_ridges", "num_peaks", "num_edges"):
        assert getattr(host, a) == getattr(dev, a), a
    for a in ("face_ridges", "ridge_peaks"):  # bit-exact, including the storage order inside a column
        h, d = getattr(host, a), getattr(dev, a)
        assert h.shape == d.shape
        assert np.array_equal(h.indptr, d.indptr) and np.array_equal(h.indices, d.indices), a
        assert np.array_equal(np.asarray(h.data, dtype=int), np.asarray(d.data, dtype=int)), a
    for a in GEOM:
        h, d = getattr(host, a), getattr(dev, a)
        assert h.shape == d.shape, a
        assert np.allclose(h, d, rtol=1e-14, atol=1e-15), a  # tolerance: the device contracts a*b+c into FMAs
    assert abs(host.mesh_size - dev.mesh_size) <= 1e-14 * host.mesh_size
    for t in ("domain_boundary_faces", "domain_boundary_nodes", "domain_boundary_ridges", "tip_ridges", "tip_peaks"):
        assert np.array_equal(host.tags[t], dev.tags[t]), t
    # cochain property of the device-built incidence matrices (tests/numerics/test_differentials.py:13-22)
    if sd.dim == 3:
        assert abs(dev.cell_faces.T.astype(int) @ dev.face_ridges.T).sum() == 0
        assert abs(dev.face_ridges.T @ dev.ridge_peaks.T).sum() == 0


def test_structured_numbering_is_the_lexicographic_one():
    """pgb_ridges3d on the raw arrays of a Kuhn grid reproduces the generator's own ridge ids."""
    ref = pg.structured_grid(3, 6)
    ref.compute_connectivity()
    dev = _raw_copy(ref)
    dev.compute_connectivity(device="cuda:0")
    assert dev.num_ridges == ref.num_ridges == 7 * 6**3 + 9 * 6**2 + 3 * 6
    assert np.array_equal(dev.ridge_peaks.indices, ref.ridge_peaks.indices)
    assert np.array_equal(dev.face_ridges.indices, ref.face_ridges.indices)
    assert np.array_equal(dev.face_ridges.data, ref.face_ridges.data)


def test_larger_ridge_sort_many_tiles():
    """> one radix tile per pass and node ids above 2^16 (several digit passes of the key)."""
    ref = pg.structured_grid(3, 40)
    ref.compute_connectivity()
    dev = _raw_copy(ref)
    dev.compute_connectivity(device="cuda:0")
    assert np.array_equal(dev.ridge_peaks.indices, ref.ridge_peaks.indices)
    assert np.array_equal(dev.face_ridges.indices, ref.face_ridges.indices)
    assert np.array_equal(dev.face_ridges.data, ref.face_ridges.data)
    assert np.array_equal(dev.tags["domain_boundary_ridges"], ref.tags["domain_boundary_ridges"])


@pytest.mark.parametrize("space", ["Nedelec0", "Lagrange2"])
def test_assembly_on_device_built_connectivity(space):
    """The engine consumes the device-resident ridges directly (no upload) and assembles the same
    matrix, bit for bit, as on the host-built connectivity."""
    sd = delaunay_grid(3, 300, seed=8)
    host, dev = _raw_copy(sd), _raw_copy(sd)
    host.compute_geometry()
    dev.compute_geometry(device="cuda:0")
    assert dev._device_ridges[2].shape[0] == host.num_ridges
    disc = getattr(pg, space)("key")
    form = "assemble_mass_matrix" if space == "Nedelec0" else "assemble_stiff_matrix"
    A, B = getattr(disc, form)(host), getattr(disc, form)(dev)
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert np.array_equal(A.data, B.data)


def test_errors():
    from pygeon_b200._lib import PgbError

    sd = delaunay_grid(3, 60, seed=1)
    bad = _raw_copy(sd)
    bad.cell_faces.data[5] *= -1  # a sign that contradicts the node-order normal
    with pytest.raises(ValueError, match="inconsistent"):
        bad.compute_geometry(device="cuda:0")
    deg = _raw_copy(sd)
    deg.face_nodes.indices[4] = deg.face_nodes.indices[3]  # a face that repeats a node
    deg.tags["domain_boundary_faces"] = np.zeros(deg.num_faces, dtype=bool)
    with pytest.raises(PgbError, match="repeats a node"):
        deg._compute_ridges_device("cuda:0")
    quad = sps.csc_array(np.ones((4, 1), dtype=bool))  # one quadrilateral face
    g = SimplexGrid(3, np.zeros((3, 4)), quad, sps.csc_array(np.ones((1, 1))), "quad")
    g.tags["domain_boundary_faces"] = np.ones(1, dtype=bool)
    with pytest.raises(NotImplementedError):
        g._compute_ridges_device("cuda:0")
