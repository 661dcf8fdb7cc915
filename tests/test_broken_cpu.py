"""Broken polynomial spaces and the inclusions into them (host side, no GPU) against the unmodified
reference: proj_to_PwPolynomials of Lagrange1 / RT0 / BDM1 / Nedelec0, PwLinears / VecPwLinears
operators, eval_at_cell_centers (fem/h1.py:218-253, hdiv.py:255-274,450-509, hcurl.py:44-61,215-291,
l2.py:435-643, vec_l2.py:127-285, discretization.py:200-219), on flat and tilted grids."""

import numpy as np
import pytest
import scipy.sparse as sps

import pygeon_b200 as pg
from oracle import ref_loader
from oracle.gen_golden import tilt
from pygeon_b200 import broken

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference source not present")
CASES = [(2, 3, False), (3, 2, False), (2, 4, True)]


def same(a, b, tol=1e-14):
    a, b = sps.csr_array(a), sps.csr_array(b)
    if a.shape != b.shape:
        return False
    if 0 in a.shape or (a.nnz == 0 and b.nnz == 0):
        return True
    return abs(a - b).max() <= tol * max(abs(b).max(), 1.0)


@pytest.fixture(scope="module", params=CASES)
def grids(request):
    dim, n, tilted = request.param
    sd = pg.structured_grid(dim, n, perturb=0.2)
    if tilted:
        tilt(sd)
    sd.compute_geometry()
    g = ref_loader.ref_grid(sd)
    sd.face_ridges = g.face_ridges  # 2D: the reference's own storage order defines the edge orientation
    return sd, g


def test_inclusions_match_reference(grids):
    sd, g = grids
    rpg, _ = ref_loader.load()
    for ours, ref in ((pg.Lagrange1(), rpg.Lagrange1()), (pg.RT0(), rpg.RT0()), (pg.BDM1(), rpg.BDM1()),
                      (pg.Nedelec0(), rpg.Nedelec0()), (pg.PwConstants(), rpg.PwConstants())):
        assert same(ours.proj_to_PwPolynomials(sd), ref.proj_to_PwPolynomials(g)), type(ours).__name__
        assert same(ours.eval_at_cell_centers(sd), ref.eval_at_cell_centers(g)), type(ours).__name__
    for ours, ref in ((pg.Lagrange1(), rpg.Lagrange1()), (pg.RT0(), rpg.RT0()), (pg.Nedelec0(), rpg.Nedelec0())):
        assert same(ours.assemble_broken_grad_matrix(sd), ref.assemble_broken_grad_matrix(g)), type(ours).__name__
    assert same(pg.eval_at_cell_centers(pg.as_mdg(sd), pg.RT0()), pg.RT0().eval_at_cell_centers(sd))


def test_pwlinears_operators_match_reference(grids):
    sd, g = grids
    rpg, _ = ref_loader.load()
    p1, r1 = broken.PwLinears(), rpg.PwLinears()
    for name in ("eval_at_cell_centers", "proj_to_lower_PwPolynomials", "get_dof_lookup_array",
                 "assemble_broken_grad_matrix", "proj_to_PwPolynomials", "assemble_diff_matrix", "assemble_stiff_matrix"):
        assert same(getattr(p1, name)(sd), getattr(r1, name)(g)), name
    assert p1.ndof(sd) == r1.ndof(g) and np.array_equal(p1.local_dofs_of_cell(sd, 1), r1.local_dofs_of_cell(g, 1))
    assert np.array_equal(p1.assemble_local_mass(sd.dim), r1.assemble_local_mass(g.dim))
    f = lambda x: np.sin(x[0]) + x[1] * x[2] + 2.0  # noqa: E731
    assert np.allclose(p1.interpolate(sd, f), r1.interpolate(g, f), rtol=1e-14, atol=1e-15)
    v1, rv = broken.VecPwLinears(), rpg.VecPwLinears()
    fv = lambda x: np.array([x[0] + 1.0, x[1] * x[0], x[2] - x[0]])  # noqa: E731
    assert np.allclose(v1.interpolate(sd, fv), rv.interpolate(g, fv), rtol=1e-14, atol=1e-15)
    assert same(v1.eval_at_cell_centers(sd), rv.eval_at_cell_centers(g))
    assert same(v1.assemble_broken_grad_matrix(sd), rv.assemble_broken_grad_matrix(g))
    assert same(broken.VecPwConstants().eval_at_cell_centers(sd), rpg.VecPwConstants().eval_at_cell_centers(g))
    assert broken.get_PwPolynomials(1, pg.VECTOR) is broken.VecPwLinears
    with pytest.raises(KeyError):
        broken.get_PwPolynomials(3, pg.SCALAR)
    assert same(broken.proj_to_PwPolynomials(pg.PwConstants(), sd, 1), rpg.proj_to_PwPolynomials(rpg.PwConstants(), g, 1))
    assert same(broken.proj_to_PwPolynomials(pg.Lagrange1(), sd, 0), rpg.proj_to_PwPolynomials(rpg.Lagrange1(), g, 0))
    assert pg.RT1().get_range_discr_class(sd.dim) is broken.PwLinears


def test_weighting_matrix_matches_reference(grids):
    sd, g = grids
    rpg, pp = ref_loader.load()
    rng = np.random.default_rng(0)
    K = rng.normal(size=(3, 3, sd.num_cells))
    sot = pp.SecondOrderTensor(np.ones(sd.num_cells))
    sot.values = K
    ref = rpg.VecPwLinears().assemble_weighting_matrix(g, sot)
    assert same(broken.VecPwLinears().assemble_weighting_matrix(sd, K), ref, 1e-13)
