"""Pin the oracle (both restatements) and the grid container against the UNMODIFIED reference,
imported from /root/reference with the porepy stand-in.  Only runs where the reference source is
present (the build container); the GPU box uses the golden fixtures generated from it."""

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import fem_oracle as fo
from oracle import ref_loader
from pygeon_b200.grid import reference_element, structured_grid

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference source not present")

TOL = 1e-12
GRIDS = [(2, 4, 0.2), (3, 3, 0.2), (3, 4, 0.0), (2, 8, 0.0)]


@pytest.mark.parametrize("mode", ["default", "sym", "nonsym"])
def test_oracle_matches_reference_on_tilted_grid(ref, mode):
    """Non-identity rotation_matrix (grids/grid.py:355-370; used at vec_l2.py:113-114, hdiv.py:493,
    hcurl.py:276): a 2D grid moved rigidly into a generic plane of R^3, on the reference pg.Grid."""
    from oracle.gen_golden import tilt

    pg, pp = ref
    sd = tilt(structured_grid(2, 4, perturb=0.2))
    sd.compute_geometry()
    g = ref_loader.ref_grid(sd)
    assert np.abs(g.rotation_matrix - np.eye(2, 3)).max() > 0.1
    assert np.allclose(g.rotation_matrix, sd.rotation_matrix, atol=1e-15)
    data = None if mode == "default" else _data(pg, pp, sd.num_cells, mode == "sym")
    cls = {"lagrange1": pg.Lagrange1, "rt0": pg.RT0, "bdm1": pg.BDM1, "nedelec0": pg.Nedelec0}
    for G in (g, sd):
        tabs = fo.cell_tables(G)
        for space, c in cls.items():
            for form, fn in (("mass", "assemble_mass_matrix"), ("stiff", "assemble_stiff_matrix")):
                refm = sps.csr_array(getattr(c("k"), fn)(g, data))
                closed = fo.closed_form(space, form, G, data, "k", tabs)
                refS = fo.canonicalise(refm, closed)
                assert np.abs(refS.data - closed.data).max() <= TOL * abs(refm).max()


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def _data(pg, pp, nc, sym, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(3, 3, nc))
    K = np.einsum("iac,jac->ijc", A, A) + np.eye(3)[:, :, None]
    if not sym:
        K = K + 0.3 * rng.normal(size=(3, 3, nc))
    sot = pp.SecondOrderTensor(np.ones(nc))
    sot.values = K
    return {pp.PARAMETERS: {"k": {pg.SECOND_ORDER_TENSOR: sot, pg.WEIGHT: rng.uniform(0.5, 2, nc)}}}


@pytest.mark.parametrize("dim,n,pert", GRIDS)
def test_grid_container_matches_reference(ref, dim, n, pert):
    """SimplexGrid computes the same ridges / geometry as pg.Grid on the same raw arrays."""
    sd = structured_grid(dim, n, perturb=pert)
    sd.compute_geometry()
    g = ref_loader.ref_grid(sd)
    for a in ("face_ridges", "ridge_peaks"):
        A, B = getattr(sd, a), getattr(g, a)
        assert A.shape == B.shape and abs(A - B).nnz == 0
    for a in ("cell_volumes", "face_normals", "cell_centers", "face_centers", "edge_tangents"):
        assert np.allclose(getattr(sd, a), getattr(g, a), rtol=1e-14, atol=1e-15)
    assert sd.num_edges == g.num_edges and abs(sd.mesh_size - g.mesh_size) < 1e-15
    assert np.array_equal(sd.compute_opposite_nodes().toarray(), g.compute_opposite_nodes().toarray())
    if dim == 3:  # cochain property, tests/numerics/test_differentials.py:13-22
        assert abs(sd.cell_faces.T.astype(int) @ sd.face_ridges.T).sum() == 0
        assert abs(sd.face_ridges.T @ sd.ridge_peaks.T).sum() == 0


@pytest.mark.parametrize("dim,n,pert", GRIDS)
@pytest.mark.parametrize("mode", ["default", "sym", "nonsym"])
@pytest.mark.parametrize("which_grid", ["reference_object", "simplex_grid"])
def test_oracle_matches_reference(ref, dim, n, pert, mode, which_grid):
    pg, pp = ref
    sd = structured_grid(dim, n, perturb=pert)
    sd.compute_geometry()
    g = ref_loader.ref_grid(sd)
    G = g if which_grid == "reference_object" else sd
    data = None if mode == "default" else _data(pg, pp, sd.num_cells, mode == "sym")
    tabs = fo.cell_tables(G)
    cls = {"lagrange1": pg.Lagrange1, "rt0": pg.RT0, "bdm1": pg.BDM1, "nedelec0": pg.Nedelec0}

    def check(refm, port, closed):
        refm = sps.csr_array(refm)
        scale = abs(refm).max()
        if port is not None:
            assert abs(refm - port).max() <= TOL * scale
            assert sps.csr_array(port).nnz == refm.nnz
        refS = fo.canonicalise(refm, closed)  # asserts pattern(ref) is inside the structural pattern
        assert np.abs(refS.data - closed.data).max() <= TOL * scale

    for space, c in cls.items():
        check(c("k").assemble_mass_matrix(g, data), fo.port_mass(space, G, data, "k", tabs),
              fo.closed_form(space, "mass", G, data, "k", tabs))
        check(c("k").assemble_stiff_matrix(g, data), fo.port_stiff(space, G, data, "k", tabs),
              fo.closed_form(space, "stiff", G, data, "k", tabs))
    check(pg.Lagrange1("k").assemble_grad_grad_matrix(g, data), None,
          fo.closed_form("lagrange1", "stiff", G, data, "k", tabs, rot2d=False))
    # Lagrange2: the reference loops over the cells in Python (h1.py:545-599)
    for form in ("mass", "stiff"):
        refm = getattr(pg.Lagrange2("k"), f"assemble_{form}_matrix")(g, data)
        check(refm, None, fo.closed_form("lagrange2", form, G, data, "k", tabs))
    # RT1: per-cell Python loop in the reference (hdiv.py:555-621)
    check(pg.RT1("k").assemble_mass_matrix(g, data), None, fo.closed_form("rt1", "mass", G, data, "k", tabs))
    d0 = pg.PwConstants("k").assemble_mass_matrix(g, data).diagonal()
    assert np.allclose(fo.p0_mass(G, data, "k").diagonal(), d0, rtol=1e-14)


@pytest.mark.parametrize("dim", [2, 3])
def test_reference_element_matches(ref, dim):
    pg, _ = ref
    a = reference_element(dim)
    b = pg.reference_element(dim)
    assert np.array_equal(a.nodes, b.nodes)
    assert np.array_equal(a.face_nodes.indices, b.face_nodes.indices)
    assert np.array_equal(a.cell_faces.toarray(), b.cell_faces.toarray())


@pytest.mark.parametrize("dim,n,pert", [(2, 5, 0.2), (3, 3, 0.2)])
def test_vectors_match_reference(ref, dim, n, pert):
    """interpolate / assemble_nat_bc (host-side vector assemblies around the path)
    against the reference's per-face Python loops (fem/hdiv.py:194-241,373-436, fem/h1.py:255-302,
    fem/l2.py:390-406, fem/hcurl.py:146-164)."""
    import pygeon_b200 as pgb

    pg, pp = ref
    sd = structured_grid(dim, n, perturb=pert)
    sd.compute_geometry()
    g = ref_loader.ref_grid(sd)
    scal = lambda x: 1.0 + x[0] - 2.0 * x[1] + 0.5 * x[2] + x[0] * x[1]  # noqa: E731
    vec = lambda x: np.array([x[1] + 1.0, -x[0] * x[2], 0.3 + x[2] if dim == 3 else 0.0])  # noqa: E731
    b_faces = sd.tags["domain_boundary_faces"]
    assert np.array_equal(b_faces, g.tags["domain_boundary_faces"])
    idx = np.where(b_faces)[0][::-1].copy()  # also an index array, in another order
    l2o, l2r = pgb.Lagrange2("k"), pg.Lagrange2("k")
    assert l2o.ndof(sd) == l2r.ndof(g)
    assert np.allclose(l2o.interpolate(sd, scal), l2r.interpolate(g, scal), rtol=1e-14, atol=1e-15)
    assert abs(l2o.assemble_diff_matrix(sd) - l2r.assemble_diff_matrix(g)).max() == 0
    rt1o, rt1r = pgb.RT1("k"), pg.RT1("k")
    assert rt1o.ndof(sd) == rt1r.ndof(g)
    assert abs(rt1o.assemble_diff_matrix(sd) - rt1r.assemble_diff_matrix(g)).max() <= 1e-13 * n
    for name, f_int in (("PwConstants", scal), ("Lagrange1", scal), ("RT0", vec), ("BDM1", vec), ("RT1", vec), ("Nedelec0", vec)):
        ours, theirs = getattr(pgb, name)("k"), getattr(pg, name)("k")
        if name == "Nedelec0" and dim == 2:
            continue  # the reference's interpolate reads ridge_peaks, which is empty in 2D (hcurl.py:160)
        assert np.allclose(ours.interpolate(sd, f_int), theirs.interpolate(g, f_int), rtol=1e-14, atol=1e-15), name
        if name == "Nedelec0":
            with pytest.raises(NotImplementedError):
                ours.assemble_nat_bc(sd, scal, b_faces)
            continue
        for bf in (b_faces, idx):
            a, b = ours.assemble_nat_bc(sd, scal, bf), theirs.assemble_nat_bc(g, scal, bf)
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-14, atol=1e-15), name


@pytest.mark.parametrize("dim,npts", [(2, 120), (3, 60)])
@pytest.mark.parametrize("mode", ["sym", "nonsym"])
def test_oracle_matches_reference_unstructured(ref, dim, npts, mode):
    """Delaunay meshes (varying node valence, faces / ridges in no lattice order): the closed forms,
    including the Lagrange2 edge-dof index rule (h1.py:529-540), against the reference itself."""
    from pgb_helpers import delaunay_grid

    pg, pp = ref
    sd = delaunay_grid(dim, npts, seed=3)
    sd.compute_geometry()
    g = ref_loader.ref_grid(sd)
    data = _data(pg, pp, sd.num_cells, mode == "sym")
    tabs = fo.cell_tables(sd)
    cls = {"lagrange1": pg.Lagrange1, "rt0": pg.RT0, "bdm1": pg.BDM1, "nedelec0": pg.Nedelec0,
           "lagrange2": pg.Lagrange2}
    for space, c in cls.items():
        for form in ("mass", "stiff"):
            refm = sps.csr_array(getattr(c("k"), f"assemble_{form}_matrix")(g, data))
            closed = fo.closed_form(space, form, sd, data, "k", tabs)
            refS = fo.canonicalise(refm, closed)
            assert np.abs(refS.data - closed.data).max() <= TOL * abs(refm).max(), (space, form)
