"""RT1 through the CUDA path: ``pg.RT1.assemble_mass_matrix`` (kind
``PGB_RT1_MASS``, 8 x 8 / 15 x 15 local blocks) against the reference's own output (golden fixtures
and the known answers of ``tests/discretizations/fem/hdiv/test_rt1.py:19-74``) and the closed-form
oracle; the host-assembled divergence / lumped matrix against the golden fixtures.

Bars as for the other spaces: indptr / indices bit-exact on the structural pattern, values within
1e-12 * max|A|, run-to-run bit-identical."""

import numpy as np
import pytest
import scipy.sparse as sps

import pygeon_b200 as pg
from oracle import fem_oracle as fo
from pgb_helpers import delaunay_grid, random_data

pytestmark = pytest.mark.gpu

TOL = 1e-12


def assert_parity(A, ref):
    """A: csc_array from the GPU path; ref: csr_array on the structural pattern."""
    assert isinstance(A, sps.csc_array)
    refc = ref.tocsc()
    refc.sort_indices()
    assert A.shape == ref.shape
    assert A.indptr.dtype == np.int32 and A.indices.dtype == np.int32
    assert np.array_equal(A.indptr, refc.indptr)
    assert np.array_equal(A.indices, refc.indices)
    assert np.abs(A.data - refc.data).max() <= TOL * np.abs(refc.data).max()


def test_golden_mass(golden):
    A = pg.RT1("key").assemble_mass_matrix(golden.sd, golden.data)
    assert_parity(A, golden.matrix("rt1", "mass"))


def _golden_general(golden, name):
    z = golden.z
    return sps.csr_array((z[name + "_data"], z[name + "_indices"], z[name + "_indptr"]), shape=tuple(z[name + "_shape"]))


def test_golden_diff_and_lumped(golden):
    rt1 = pg.RT1("key")
    ref = _golden_general(golden, "rt1_diff")
    D = rt1.assemble_diff_matrix(golden.sd)
    assert D.shape == ref.shape and abs(sps.csr_array(D) - ref).max() <= TOL * abs(ref).max()
    ref = _golden_general(golden, "rt1_lumped")
    Lm = rt1.assemble_lumped_matrix(golden.sd, golden.data)
    assert Lm.shape == ref.shape and abs(sps.csr_array(Lm) - ref).max() <= TOL * abs(ref).max()


@pytest.mark.parametrize("dim,n", [(2, 37), (3, 9)])
@pytest.mark.parametrize("mode", ["default", "sym", "nonsym"])
def test_oracle_parity_medium(dim, n, mode):
    sd = pg.structured_grid(dim, n, perturb=0.2, seed=3)
    sd.compute_geometry()
    data = None if mode == "default" else random_data(sd.num_cells, seed=5, sym=(mode == "sym"))
    A = pg.RT1("key").assemble_mass_matrix(sd, data)
    assert_parity(A, fo.closed_form("rt1", "mass", sd, data, "key"))


@pytest.mark.parametrize("dim,npts", [(2, 3000), (3, 1500)])
def test_oracle_parity_unstructured(dim, npts):
    sd = delaunay_grid(dim, npts, seed=5)
    sd.compute_geometry()
    data = random_data(sd.num_cells, seed=7, sym=False)
    A = pg.RT1("key").assemble_mass_matrix(sd, data)
    assert_parity(A, fo.closed_form("rt1", "mass", sd, data, "key"))


def test_full_sort_fallback_agrees():
    """mode 1 (64-bit radix sort of all COO keys) gives the same pattern and values as the row kernels."""
    from pygeon_b200 import engine

    sd = pg.structured_grid(3, 5, perturb=0.2)
    sd.compute_geometry()
    pack = engine.get_pack(sd)
    outs = []
    for mode in (0, 1):
        pack.plans.clear()
        plan = engine.symbolic(pack, "rt1", mode=mode)
        data = engine.numeric(plan, engine.local_matrices(pack, "rt1_mass", plan=plan))
        outs.append((plan.indptr.cpu().numpy(), plan.indices.cpu().numpy(), data.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.abs(outs[0][2] - outs[1][2]).max() <= 1e-14 * np.abs(outs[0][2]).max()


def test_bitwise_reproducible_and_properties():
    from pygeon_b200 import engine

    sd = pg.structured_grid(3, 8, perturb=0.2)
    sd.compute_geometry()
    data = random_data(sd.num_cells, seed=1)
    outs = []
    for _ in range(3):
        engine.clear_cache(sd)
        A = pg.RT1("key").assemble_mass_matrix(sd, data)
        outs.append((A.indptr.copy(), A.indices.copy(), A.data.copy()))
    for o in outs[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(outs[0], o))
    # symmetric positive definite for a symmetric tensor; reproduces constants:
    # u = interpolate(const) gives u^T M u = |const|^2 |Omega| for K = I
    M = pg.RT1().assemble_mass_matrix(sd)
    assert abs(M - M.T).max() <= 1e-14 * abs(M).max()
    u = pg.RT1().interpolate(sd, lambda x: np.array([1.0, -2.0, 0.5]))
    assert abs(u @ (M @ u) - 5.25) < 1e-11
    assert pg.RT1().ndof(sd) == 3 * (sd.num_faces + sd.num_cells)
    # stiffness (round 2): symmetric positive semi-definite, kernel contains the divergence-free interpolant
    S = pg.RT1().assemble_stiff_matrix(sd)
    assert abs(S - S.T).max() <= 1e-12 * abs(S).max() and abs(u @ (S @ u)) < 1e-9


# ---- round 2: RT1 stiffness (range space PwLinears), broken-P1 mass, error_l2 ---------------------------
def _csr(golden, key, n):
    z = golden.z
    return sps.csr_array((z[key + "_data"], z[key + "_indices"], z[key + "_indptr"]), shape=(n, n))


def test_rt1_stiffness_golden(golden):
    """RT1.assemble_stiff_matrix = div^T M_PwLinears div (discretization.py:154-183, hdiv.py:813-865,958)
    through K1 kind PGB_RT1_DIVDIV, against the reference's matrix on the structural pattern."""
    sd = golden.sd
    A = pg.RT1("key").assemble_stiff_matrix(sd, golden.data)
    n = pg.RT1().ndof(sd)
    ref = _csr(golden, "rt1_stiff", n).tocsc()
    ref.sort_indices()
    assert A.shape == (n, n)
    assert np.array_equal(A.indptr, ref.indptr) and np.array_equal(A.indices, ref.indices)
    assert np.abs(A.data - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
    # and equals B^T M B with this package's own divergence and broken-P1 mass
    B = pg.RT1("key").assemble_diff_matrix(sd)
    M = pg.PwLinears("key").assemble_mass_matrix(sd, golden.data)
    assert abs(A - B.T @ M @ B).max() <= 1e-12 * abs(A).max()


def test_pwlinears_mass_and_lumped_golden(golden):
    sd = golden.sd
    n = (sd.dim + 1) * sd.num_cells
    for nm, A in (("pwlinears_mass", pg.PwLinears("key").assemble_mass_matrix(sd, golden.data)),
                  ("pwlinears_lumped", pg.PwLinears("key").assemble_lumped_matrix(sd, golden.data))):
        ref = _csr(golden, nm, n).tocsc()
        ref.sort_indices()
        assert isinstance(A, sps.csc_array) and A.shape == (n, n)
        assert np.array_equal(A.indptr, ref.indptr) and np.array_equal(A.indices, ref.indices), nm
        assert np.abs(A.data - ref.data).max() <= 1e-13 * np.abs(ref.data).max(), nm


def test_error_l2_golden(golden):
    """Discretization.error_l2 (discretization.py:304-358): default (compared in the broken P1 space),
    absolute, and against the space's own interpolant."""
    sd = golden.sd
    fs = lambda x: np.sin(x[0]) + x[1] * x[2] + 2.0  # noqa: E731
    fv = lambda x: np.array([x[0] + 1.0, x[1] * x[0], x[2] - x[0]])  # noqa: E731
    for nm, d, f in (("lagrange1", pg.Lagrange1("key"), fs), ("rt0", pg.RT0("key"), fv), ("bdm1", pg.BDM1("key"), fv),
                     ("nedelec0", pg.Nedelec0("key"), fv), ("p0", pg.PwConstants("key"), fs)):
        if "errl2_" + nm not in golden.z.files:
            continue
        u, ref = golden.z["errl2_u_" + nm], golden.z["errl2_" + nm]
        got = [d.error_l2(sd, u, f, data=golden.data), d.error_l2(sd, u, f, relative=False, data=golden.data),
               d.error_l2(sd, u, f, poly_order=None, data=golden.data)]
        assert np.allclose(got, ref, rtol=1e-10), (nm, got, ref)
