"""Block operators in HBM: ``pgb_csr_transpose`` (scipy's ``A.T.tocsr()``), ``engine.device_bmat``
(``sps.block_array``, reference utils/bmat.py + numerics/innerproducts.py:162-232), ``pgb_csr_add_diagonal``
(the mortar trace term) and the device-resident mixed-dimensional Darcy saddle system built from them
(``pygeon_b200/md_system.py``) against the host operators, which tests/test_mdg_gpu.py pins to the reference.

Integer arrays bit-exact; values exactly equal (copies, one scaling by -1, one addition)."""

import numpy as np
import pytest
import scipy.sparse as sps
import torch

import pygeon_b200 as pg
from pygeon_b200 import engine
from pygeon_b200._lib import PgbError

pytestmark = pytest.mark.gpu


def rand_csr(rng, m, n, per_row, empty_every=0):
    """Random CSR with sorted distinct columns; every `empty_every`-th row empty."""
    rows, cols = [], []
    for r in range(m):
        if empty_every and r % empty_every == 0:
            continue
        c = np.unique(rng.integers(0, n, size=int(rng.integers(0, per_row + 1))))
        rows.append(np.full(c.size, r))
        cols.append(c)
    rows = np.concatenate(rows) if rows else np.zeros(0, int)
    cols = np.concatenate(cols) if cols else np.zeros(0, int)
    A = sps.csr_array((rng.standard_normal(rows.size), (rows, cols)), shape=(m, n))
    A.sort_indices()
    return A


def same_csr(D: engine.DeviceCSR, ref):
    ref = sps.csr_array(ref)
    ref.sort_indices()
    got = D.to_scipy("csr")  # the arrays as they are
    assert got.shape == ref.shape
    assert np.array_equal(got.indptr, ref.indptr)
    assert np.array_equal(got.indices, ref.indices)
    assert np.array_equal(got.data, ref.data)


@pytest.mark.parametrize("m,n,per_row,empty", [(1, 1, 1, 0), (50, 7, 5, 3), (300, 70001, 9, 4), (4000, 1000, 40, 0),
                                               (17, 20_000_000, 6, 0), (5, 3, 0, 0)])
def test_transpose_matches_scipy(m, n, per_row, empty):
    rng = np.random.default_rng(m + n)
    A = rand_csr(rng, m, n, per_row, empty)
    D = engine.DeviceCSR.from_scipy(A)
    T = D.transpose()
    assert T.shape == (n, m) and not T.symmetric
    same_csr(T, A.T.tocsr())
    same_csr(T.transpose(), A)  # involution
    # deterministic: bit-identical on a second run
    T2 = D.transpose()
    assert torch.equal(T.indices, T2.indices) and torch.equal(T.data, T2.data) and torch.equal(T.indptr, T2.indptr)


def test_transpose_int64_indptr(monkeypatch):
    rng = np.random.default_rng(5)
    A = rand_csr(rng, 600, 431, 12, 5)
    monkeypatch.setattr(engine, "force_idx64", True)
    D = engine.DeviceCSR.from_scipy(A)
    D.indptr = D.indptr.to(torch.int64)
    T = D.transpose()
    assert T.indptr.dtype == torch.int64
    same_csr(T, A.T.tocsr())


def test_device_bmat_matches_block_array():
    rng = np.random.default_rng(11)
    hs, ws = [40, 1, 25], [30, 12, 7, 3]
    host = [[None] * len(ws) for _ in hs]
    dev = [[None] * len(ws) for _ in hs]
    scale = {(1, 2): -1.0, (2, 3): 2.5}
    for i, h in enumerate(hs):
        for j, w in enumerate(ws):
            if (i + 2 * j) % 3 == 1 and not (i == 1 and j == 0):
                continue
            A = rand_csr(rng, h, w, 6, 4 if (i + j) % 2 else 0)
            s = scale.get((i, j), 1.0)
            host[i][j] = s * A
            D = engine.DeviceCSR.from_scipy(A)
            if (i + j) % 2:  # int64 row pointers in some blocks
                D.indptr = D.indptr.to(torch.int64)
            dev[i][j] = (D, s) if (i, j) in scale else D
    for j in range(len(ws)):  # every block column needs a block to be sized
        assert any(host[i][j] is not None for i in range(len(hs)))
    out = engine.device_bmat(dev)
    assert out.shape == (sum(hs), sum(ws))
    same_csr(out, sps.block_array(host).tocsr())
    with pytest.raises(ValueError):
        engine.device_bmat([[dev[0][0], None], [None, None]])
    with pytest.raises(ValueError):
        engine.device_bmat([[dev[0][0]], [dev[1][1]]])  # widths 30 and 12 in one block column


def test_add_diagonal():
    rng = np.random.default_rng(2)
    A = rand_csr(rng, 200, 200, 8) + sps.eye_array(200, format="csr")
    A = sps.csr_array(A)
    A.sort_indices()
    D = engine.DeviceCSR.from_scipy(A)
    rows = rng.choice(200, size=37, replace=False)
    vals = rng.standard_normal(37)
    D.add_diagonal(torch.from_numpy(rows.astype(np.int32)), torch.from_numpy(vals))
    ref = A.copy().tolil()
    for r, v in zip(rows, vals):
        ref[r, r] = A[r, r] + v
    same_csr(D, ref.tocsr())
    B = sps.csr_array(([1.0], ([0], [1])), shape=(2, 2))
    with pytest.raises(PgbError):
        engine.DeviceCSR.from_scipy(B).add_diagonal(torch.tensor([0], dtype=torch.int32), torch.tensor([1.0], dtype=torch.float64))


@pytest.fixture(scope="module", params=[(2, 0.0), (6, 0.2)])
def mdg(request):
    n, perturb = request.param
    g = pg.fractured_unit_cube(n, perturb=perturb, seed=4)
    g.compute_geometry()
    rng = np.random.default_rng(n)
    for sd, d_sd in g.subdomains(return_data=True):
        k = rng.uniform(0.5, 2.0, sd.num_cells)
        d_sd.update({pg.PARAMETERS: {"flow": {pg.SECOND_ORDER_TENSOR: pg.SecondOrderTensor(k), pg.WEIGHT: rng.uniform(0.5, 2.0, sd.num_cells)}}})
    intf, d_intf = g.interfaces(return_data=True)[0]
    d_intf.update({pg.PARAMETERS: {"flow": {pg.NORMAL_DIFFUSIVITY: rng.uniform(1.0, 50.0, intf.num_cells)}}})
    return g


def test_face_mass_device_equals_host_operator(mdg):
    # The engine's CSR arrays are the CSC arrays of the reference's matrix (pgb.h, K1 note): a[i][j] and a[j][i]
    # of a symmetric form are computed separately and may differ in the last bit, so the exact comparison is
    # with the transposed host matrix.
    ref = pg.face_mass(mdg, keyword="flow")
    M = pg.face_mass_device(mdg, keyword="flow")
    assert M.symmetric and isinstance(ref, sps.csc_array)
    same_csr(M, ref.T)
    assert abs(sps.csr_array(ref) - M.to_scipy("csr")).max() <= 1e-15 * abs(ref).max()
    same_csr(pg.face_mass_device(mdg, keyword="flow", trace_contribution=False),
             pg.face_mass(mdg, keyword="flow", trace_contribution=False).T)


def test_saddle_system_on_the_device(mdg):
    """The stacked matrix equals the host stack of the host operators entry for entry; the device MINRES with
    the block-Jacobi diagonal made in HBM reproduces the host-path solve."""
    M = pg.face_mass(mdg, keyword="flow")
    B = sps.csr_array(pg.cell_mass(mdg, keyword="flow") @ pg.div(mdg))
    ref = sps.block_array([[M, -B.T], [-B, None]]).tocsr()
    A, nf, nc = pg.mdg_darcy_saddle_device(mdg, keyword="flow")
    assert (nf, nc) == (M.shape[0], B.shape[0]) and A.shape == ref.shape
    same_csr(A, sps.block_array([[M.T, -B.T], [-B, None]]))  # M.T: see test_face_mass_device_equals_host_operator
    assert abs(ref - A.to_scipy("csr")).max() <= 1e-15 * abs(ref).max()
    sd3, sd2 = mdg.subdomains()
    rhs = np.zeros(nf + nc)
    rhs[nf + sd3.num_cells:] = -1.0
    dinv = pg.darcy_block_jacobi(A, nf)
    dinv_host = 1.0 / np.concatenate([M.diagonal(), np.asarray(B.multiply(B) @ (1.0 / M.diagonal())).ravel()])
    assert np.allclose(dinv.cpu().numpy(), dinv_host, rtol=1e-13)
    b = torch.from_numpy(rhs).cuda()
    x, info = pg.minres_device(A, b, rtol=1e-12, maxiter=20000, dinv=dinv)
    assert info.info == 0
    xh, ih = pg.minres(ref, rhs, rtol=1e-12, maxiter=20000, M=dinv_host)
    assert ih == 0
    x = x.cpu().numpy()
    assert np.abs(x - xh).max() <= 1e-8 * np.abs(xh).max()
    assert np.linalg.norm(ref @ x - rhs) <= 1e-9 * np.linalg.norm(rhs)
