"""Mixed-dimensional block operators (BASELINE config #5 in miniature) through the CUDA
path against fixtures generated from the UNMODIFIED reference by oracle/gen_golden_mdg.py: pg.cell_mass /
face_mass (with the mortar trace term) / ridge_mass / peak_mass / lumped_face_mass / face_stiff
(numerics/innerproducts.py:162-232, stiffness.py:80-109) on a cube cut by a fracture plane, and the
mixed-dimensional Darcy solve with the GPU MINRES."""

import glob
import os

import numpy as np
import pytest
import scipy.sparse as sps

import pygeon_b200 as pg

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HERE, "golden_mdg", "*.npz")))


@pytest.fixture(scope="module", params=NAMES)
def case(request):
    z = np.load(os.path.join(HERE, "golden_mdg", request.param + ".npz"))
    mdg = pg.fractured_unit_cube(int(z["n"]), perturb=float(z["perturb"]), seed=3)
    mdg.compute_geometry()
    for i, (sd, d_sd) in enumerate(mdg.subdomains(return_data=True)):
        sot = pg.SecondOrderTensor(np.ones(sd.num_cells))
        sot.values = z[f"K{i}"]
        d_sd.update({pg.PARAMETERS: {"flow": {pg.SECOND_ORDER_TENSOR: sot, pg.WEIGHT: z[f"w{i}"]}}})
    intf, d_intf = mdg.interfaces(return_data=True)[0]
    d_intf.update({pg.PARAMETERS: {"flow": {pg.NORMAL_DIFFUSIVITY: z["kn"]}}})
    return mdg, z


def golden(z, key):
    return sps.csr_array((z[key + "_data"], z[key + "_indices"], z[key + "_indptr"]), shape=tuple(z[key + "_shape"]))


@pytest.mark.parametrize("name", ["cell_mass", "face_mass", "ridge_mass", "peak_mass", "lumped_face_mass", "face_stiff"])
def test_block_operators_match_reference(case, name):
    mdg, z = case
    ref = golden(z, name)
    A = sps.csr_array(getattr(pg, name)(mdg, keyword="flow"))
    assert A.shape == ref.shape
    assert abs(A - ref).max() <= 1e-12 * abs(ref).max()
    # the reference's pattern (exact zeros pruned by scipy) lies inside the structural pattern built here
    pat = A.copy()
    pat.data = np.ones_like(pat.data)
    assert (abs(ref) > 0).multiply(pat).nnz == (abs(ref) > 0).nnz
    if not name.endswith("stiff"):  # the stiffness dispatcher has no block form (stiffness.py:80-109)
        blocks = getattr(pg, name)(mdg, keyword="flow", as_bmat=True)
        assert blocks.shape == (2, 2) and abs(sps.block_array(blocks) - A).max() == 0


def test_mixed_dimensional_darcy_solve(case):
    """[[M, -B^T], [B, 0]] with M = face_mass (mortar term included), B = cell_mass @ div, symmetrised for the GPU
    MINRES (second block row negated), against the reference system's SuperLU solution."""
    mdg, z = case
    M, B = pg.face_mass(mdg, keyword="flow"), pg.cell_mass(mdg, keyword="flow") @ pg.div(mdg)
    nf = M.shape[0]
    sym = sps.block_array([[M, -B.T], [-B, None]]).tocsr()
    rhs = z["darcy_rhs"].copy()
    rhs[nf:] *= -1.0
    dinv = 1.0 / np.concatenate([M.diagonal(), np.asarray((B.multiply(B) @ (1.0 / M.diagonal()))).ravel()])
    x, info = pg.minres(sym, rhs, rtol=1e-13, maxiter=20000, M=dinv)
    ref = z["darcy_sol"]
    assert info == 0 and np.abs(x - ref).max() <= 1e-7 * np.abs(ref).max()
    # mass conservation in every fracture cell: div q (with the jump of the matrix fluxes) = source
    nc3 = mdg.subdomains()[0].num_cells
    assert np.allclose((B @ x[:nf])[nc3:], z["darcy_rhs"][nf + nc3:], atol=1e-8)
