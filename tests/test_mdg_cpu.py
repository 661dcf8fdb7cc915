"""Mixed-dimensional grids (BASELINE config #5) on the host: the conforming mortar grid's
signed maps and jump operators, leaf tags and the block exterior derivatives against the UNMODIFIED reference
(grids/mortar_grid.py:46-208, grids/md_grid.py:113-153, numerics/differentials.py:144-192,
numerics/restrictions.py:11-51, innerproducts.py:207-224) on a cube cut by a fracture plane; cochain
property of the mixed-dimensional complex (tests/numerics/test_differentials.py:13-43)."""

import numpy as np
import pytest
import scipy.sparse as sps

import pygeon_b200 as pg
from oracle import ref_loader


def same(a, b):
    a, b = sps.csr_array(a), sps.csr_array(b)
    return a.shape == b.shape and (a.nnz == b.nnz == 0 or abs(a - b).max() <= 1e-14)


@pytest.fixture(scope="module", params=[(2, 0.0), (4, 0.15)])
def cube(request):
    n, pert = request.param
    mdg = pg.fractured_unit_cube(n, perturb=pert)
    mdg.compute_geometry()
    return mdg


def test_cochain_and_shapes(cube):
    mdg = cube
    sd3, sd2 = mdg.subdomains()
    intf = mdg.interfaces()[0]
    assert (sd3.dim, sd2.dim, intf.dim) == (3, 2, 2) and intf.num_cells == 2 * sd2.num_cells
    assert mdg.interface_to_subdomain_pair(intf) == (sd3, sd2)
    grad, curl, div = pg.grad(mdg), pg.curl(mdg), pg.div(mdg)
    assert (curl @ grad).nnz == 0 and (div @ curl).nnz == 0
    assert div.shape == (mdg.num_subdomain_cells(), mdg.num_subdomain_faces())
    assert curl.shape == (mdg.num_subdomain_faces(), mdg.num_subdomain_ridges())
    assert grad.shape == (mdg.num_subdomain_ridges(), mdg.num_subdomain_peaks())
    # every fracture cell receives the flux of its two sides: the jump block has two entries per row
    jump = pg.div(mdg, as_bmat=True)[1, 0]
    assert np.all(np.diff(sps.csr_array(jump).indptr) == 2)
    # leaf tags: the fracture faces; no ridge / peak of the 3D grid coincides with a grid two dimensions down
    assert sd3.tags["leaf_faces"].sum() == intf.num_cells
    assert sd3.tags["leaf_ridges"].sum() == 0 and sd3.tags["leaf_peaks"].sum() == 0
    assert not sd3.tags["domain_boundary_faces"][sd3.tags["fracture_faces"]].any()


@pytest.mark.skipif(not ref_loader.available(), reason="reference source not present")
def test_mortar_and_block_operators_match_reference(cube):
    from oracle.ref_mdg import ref_mdg

    rpg, _ = ref_loader.load()
    mdg = cube
    rm, table = ref_mdg(mdg)
    intf, rintf = mdg.interfaces()[0], rm.interfaces()[0]
    for name in ("signed_mortar_to_primary", "cell_faces", "face_ridges", "ridge_peaks"):
        assert same(getattr(intf, name), getattr(rintf, name)), name
    assert np.allclose(intf.cell_volumes, rintf.cell_volumes)
    for sd in mdg.subdomains():
        g = table[id(sd)]
        for tag in ("leaf_faces", "leaf_ridges", "leaf_peaks"):
            assert np.array_equal(np.asarray(sd.tags[tag], dtype=bool), np.asarray(g.tags[tag], dtype=bool)), tag
    for name in ("div", "curl", "grad"):
        assert same(getattr(pg, name)(mdg), getattr(rpg, name)(rm)), name
        ob, rb = getattr(pg, name)(mdg, as_bmat=True), getattr(rpg, name)(rm, as_bmat=True)
        assert all(same(ob[i, j], rb[i, j]) for i in range(2) for j in range(2))
    for k in range(4):
        assert same(pg.numerics.restrictions.zero_tip_dofs(mdg, k), rpg.numerics.restrictions.zero_tip_dofs(rm, k))
    assert same(pg.remove_tip_dofs(mdg, 1), rpg.remove_tip_dofs(rm, 1))
    # the mortar term of the face mass matrix in isolation (zero subdomain blocks): S diag(1 / (|m| k_n)) S^T
    kn = np.linspace(0.5, 2.0, intf.num_cells)
    mdg.interface_data(intf).update({pg.PARAMETERS: {"flow": {pg.NORMAL_DIFFUSIVITY: kn}}})
    rm.interface_data(rintf).update({"parameters": {"flow": {rpg.NORMAL_DIFFUSIVITY: kn}}})
    zero = lambda sd, n_minus_k, discr, d_sd, **kw: sps.csc_array((sd.num_faces, sd.num_faces))  # noqa: E731
    ours = pg.numerics.innerproducts.mass_matrix(mdg, 1, None, zero, keyword="flow")
    ref = rpg.numerics.innerproducts.mass_matrix(rm, 1, None, zero, keyword="flow")
    assert ours.nnz == intf.num_cells and same(ours, ref)
    assert pg.numerics.innerproducts.mass_matrix(mdg, 1, None, zero, keyword="flow", trace_contribution=False).nnz == 0


def test_trace_term_is_diagonal_on_a_conforming_interface(cube):
    """Host side of the device-resident face mass (md_system._trace_diagonal): the mortar trace term
    S diag(1 / (|m| k_n)) S^T (innerproducts.py:207-224) touches one diagonal entry per fracture face of the 3D
    grid; a non-diagonal term (non-conforming interface) is refused, not approximated."""
    from pygeon_b200 import md_system

    mdg = cube
    intf, d_intf = mdg.interfaces(return_data=True)[0]
    rng = np.random.default_rng(1)
    kn = rng.uniform(0.5, 5.0, intf.num_cells)
    d_intf.update({pg.PARAMETERS: {"flow": {pg.NORMAL_DIFFUSIVITY: kn}}})
    rows, vals = md_system._trace_diagonal(intf, d_intf, "flow")
    S = intf.signed_mortar_to_primary
    T = sps.csr_array(S @ sps.diags_array(1.0 / intf.cell_volumes / kn) @ S.T)
    assert rows.size == intf.num_cells and np.array_equal(np.unique(rows), rows)
    assert np.array_equal(vals, T.diagonal()[rows]) and abs(T - sps.diags_array(T.diagonal())).max() == 0
    sd3 = mdg.subdomains()[0]
    assert np.all(sd3.tags["fracture_faces"][rows])

    class Fake:  # two mortar cells on one primary face: the term couples them
        signed_mortar_to_primary = sps.csc_array(np.array([[1.0, 0.0], [1.0, -1.0]]))
        cell_volumes = np.ones(2)
        num_cells = 2

    with pytest.raises(NotImplementedError):
        md_system._trace_diagonal(Fake(), {pg.PARAMETERS: {"flow": {pg.NORMAL_DIFFUSIVITY: np.ones(2)}}}, "flow")


def test_left_scaling_equals_the_sparse_product():
    """md_system._left_scaled: rows of div scaled by the diagonal P0 mass == cell_mass @ div of the host route
    (same pattern - exact zeros dropped - and bit-identical values); a non-diagonal factor takes the product."""
    from pygeon_b200 import md_system

    rng = np.random.default_rng(3)
    D = sps.random_array((300, 500), density=0.02, rng=rng, format="csr")
    D.data[::7] = 0.0  # explicit zeros (masked tip dofs)
    d = rng.uniform(0.5, 2.0, 300)
    d[5] = 0.0
    for Cm in (sps.diags_array(d).tocsc(), sps.csc_array(sps.diags_array(d) + sps.eye_array(300, k=1) * 0.1)):
        got, ref = md_system._left_scaled(Cm, D), sps.csr_array(Cm @ D)
        ref.sort_indices(), got.sort_indices()  # (scipy's product has already dropped the exact zeros)
        assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
        assert np.array_equal(got.data, ref.data)
