"""Consistency identities the reference's own test-suite checks for the spaces on the path
(SURVEY §4 item 2), restated for the CPU oracle's closed forms on the structured `unit_sd`
variants.  CPU only: pins the oracle (and the host-side interpolation) further; the CUDA path is
compared with the same closed forms in the ``-m gpu`` tests."""

import numpy as np
import pytest
import scipy.sparse as sps

import pygeon_b200 as pgb
from oracle import fem_oracle as fo
from pygeon_b200.grid import structured_grid

GRIDS = [(2, 4, 0.0), (2, 5, 0.2), (3, 2, 0.0), (3, 3, 0.2)]


def _grid(dim, n, pert):
    sd = structured_grid(dim, n, perturb=pert, seed=2)
    sd.compute_geometry()
    return sd


@pytest.mark.parametrize("dim,n,pert", GRIDS)
def test_bdm1_mass_vs_rt0(dim, n, pert):
    """tests/discretizations/fem/hdiv/test_bdm1.py:75-84: M_RT0 = P^T M_BDM1 P, P = proj_from_RT0."""
    sd = _grid(dim, n, pert)
    M0 = fo.closed_form("rt0", "mass", sd)
    M1 = fo.closed_form("bdm1", "mass", sd)
    P = sps.vstack([sps.eye_array(sd.num_faces)] * dim).tocsr()
    assert abs(M0 - P.T @ M1 @ P).max() <= 1e-13 * abs(M0).max()


@pytest.mark.parametrize("dim,n,pert", GRIDS)
@pytest.mark.parametrize("space", ["bdm1", "rt1"])
def test_norm_of_linear_function(dim, n, pert, space):
    """tests/discretizations/fem/hdiv/test_hdiv.py:47-59: ||x||^2_M = d / 3 on the unit square / cube."""
    sd = _grid(dim, n, pert)
    disc = {"bdm1": pgb.BDM1, "rt1": pgb.RT1}[space]()
    u = disc.interpolate(sd, lambda x: x)
    M = fo.closed_form(space, "mass", sd)
    assert np.isclose(u @ (M @ u), dim / 3, rtol=1e-12)


@pytest.mark.parametrize("dim,n,pert", GRIDS)
def test_rt0_rt1_reproduce_constants(dim, n, pert):
    """Constant fields lie in every H(div) space on the path: u^T M u = |c|^2 |Omega| (the energy
    version of test_hdiv.py:21-30, which needs eval_at_cell_centers)."""
    sd = _grid(dim, n, pert)
    c = np.array([2.0, 3.0, -1.0 if dim == 3 else 0.0])
    for space, disc in (("rt0", pgb.RT0()), ("bdm1", pgb.BDM1()), ("rt1", pgb.RT1())):
        u = disc.interpolate(sd, lambda _: c)
        M = fo.closed_form(space, "mass", sd)
        assert np.isclose(u @ (M @ u), c @ c, rtol=1e-12), space


@pytest.mark.parametrize("dim,n,pert", GRIDS)
def test_lumped_consistency(dim, n, pert):
    """tests/discretizations/fem/h1/test_h1.py:29-38: lumped and full mass integrate the interpolant of
    1 identically (Lagrange1)."""
    sd = _grid(dim, n, pert)
    one = np.ones(sd.num_nodes)
    ML = fo.closed_form("lagrange1", "lumped", sd)
    MF = fo.closed_form("lagrange1", "mass", sd)
    assert np.allclose(ML @ one, MF @ one, rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("dim,n,pert", GRIDS)
def test_stiffness_consistency(dim, n, pert):
    """tests/discretizations/fem/h1/test_h1.py:41-53: B^T M B equals the grad-grad matrix (K = I)."""
    sd = _grid(dim, n, pert)
    S1 = fo.closed_form("lagrange1", "stiff", sd)
    S2 = fo.closed_form("lagrange1", "stiff", sd, rot2d=False)
    assert abs(S1 - S2).max() <= 1e-13 * abs(S1).max()
    S3 = fo.port_stiff("lagrange1", sd)
    assert abs(S1 - sps.csr_array(S3)).max() <= 1e-12 * abs(S1).max()


@pytest.mark.parametrize("dim,n,pert", GRIDS)
def test_rt1_divergence_of_linear_field(dim, n, pert):
    """div x = d: the RT1 divergence (host assembly, fem/hdiv.py:813-865) of the interpolant of x is the
    constant d in every PwLinears dof."""
    sd = _grid(dim, n, pert)
    rt1 = pgb.RT1()
    div = rt1.assemble_diff_matrix(sd) @ rt1.interpolate(sd, lambda x: x)
    assert div.shape == ((dim + 1) * sd.num_cells,) and np.allclose(div, dim, rtol=1e-11)
