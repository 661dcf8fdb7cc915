"""Parity at the sizes the benchmark numbers are quoted on ("what's weak" #1): the full
CSR of BASELINE config #2 (pg.Lagrange1 stiffness on 128^3 Kuhn tets: 12.6 M cells, 2.0e8 COO
candidates, uint32 positions, the per-chunk kernel selection) and of config #1 (2D mixed Darcy on
256^2: RT0 mass and P0 mass @ div) against the oracle's closed forms, bit-exact pattern and
1e-12 * max|A| values; plus a row sample of the config-#3-sized one-pass saddle matrix when the
device has room for it."""

import numpy as np
import psutil
import pytest
import scipy.sparse as sps
import torch

import pygeon_b200 as pg
from oracle import fem_oracle as fo

pytestmark = pytest.mark.gpu
TOL = 1e-12


def full_parity(A, ref):
    refc = ref.tocsc()
    refc.sort_indices()
    assert A.shape == refc.shape and A.nnz == refc.nnz
    assert np.array_equal(A.indptr, refc.indptr)
    assert np.array_equal(A.indices, refc.indices)
    err = np.abs(A.data - refc.data).max() / np.abs(refc.data).max()
    assert err <= TOL, err
    return err


def test_config2_full_csr_against_oracle():
    n = 128 if psutil.virtual_memory().available > 56 << 30 else 96
    sd = pg.structured_grid(3, n, device="cuda")
    sd.compute_connectivity()
    A = pg.Lagrange1().assemble_stiff_matrix(sd)
    assert A.shape == ((n + 1) ** 3,) * 2
    ref = fo.closed_form("lagrange1", "stiff", sd, fast=True)
    assert ref.nnz == (n + 1) ** 3 + 2 * sd.num_ridges  # nodes + 2 * edges
    full_parity(A, ref)


def test_config1_full_against_oracle_and_port():
    """256 x 256 squares: RT0 mass (structural pattern; the projection-route port canonicalised onto
    it), B = P0 mass @ div, and the block system the reference's tutorial stacks (darcy.ipynb cell 10)."""
    sd = pg.structured_grid(2, 256)
    sd.compute_geometry()
    M = pg.RT0().assemble_mass_matrix(sd)
    ref = fo.closed_form("rt0", "mass", sd)
    assert ref.nnz == 983552
    full_parity(M, ref)
    port = fo.port_mass("rt0", sd)
    portS = fo.canonicalise(port, ref)
    assert np.abs(portS.data - ref.data).max() <= TOL * np.abs(ref.data).max()
    P0 = pg.PwConstants().assemble_mass_matrix(sd)
    assert np.allclose(P0.diagonal(), 1.0 / sd.cell_volumes, rtol=1e-13)
    B = P0 @ pg.div(sd)
    Bref = sps.diags_array(1.0 / sd.cell_volumes) @ sd.cell_faces.T
    assert abs(B - Bref).max() <= 1e-13 * abs(Bref).max()
    saddle = pg.darcy_saddle_matrix(sd)
    blk = sps.block_array([[ref, -Bref.T], [-Bref, sps.csr_array((sd.num_cells, sd.num_cells))]]).tocsc()
    assert abs(saddle - blk).max() <= TOL * abs(blk).max()


def sampled_rows_reference(space, form, sd, rows, dofs):
    """Rows ``rows`` of the closed-form matrix from the cells that touch them (any grid size): the
    oracle is run on the sub-grid made of those cells (global node / face ids kept)."""
    cells = np.flatnonzero(np.isin(dofs, rows).any(axis=1))
    sub = pg.SimplexGrid(sd.dim, sd.nodes, sd.face_nodes, sd.cell_faces[:, cells], "sample")
    sub.rotation_matrix = np.eye(sd.dim, 3)
    A, d = fo.local_matrices(space, form, sub)
    r = np.broadcast_to(d[:, :, None], A.shape).ravel()
    c = np.broadcast_to(d[:, None, :], A.shape).ravel()
    sel = np.isin(r, rows)
    return fo.coo_to_csr(r[sel], c[sel], A.ravel()[sel], fo.NDOF[space](sub))


def test_int64_and_uint32_limits_by_row_sample():
    """RT0 mass at 160^3 (24.6 M tets, 3.9e8 COO candidates; int64 indptr forced): 20 000 sampled rows of
    the device CSR against the oracle evaluated on the cells that touch them."""
    from pygeon_b200 import engine

    if torch.cuda.mem_get_info()[0] < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    n = 160
    sd = pg.structured_grid(3, n, device="cuda")
    sd.compute_connectivity()
    engine.force_idx64 = True
    try:
        Ad = pg.RT0().assemble_mass_matrix_device(sd)
    finally:
        engine.force_idx64 = False
    assert Ad.indptr.dtype == torch.int64 and Ad.nnz == 16 * sd.num_cells - (sd.num_faces - int(sd.tags["domain_boundary_faces"].sum()))
    rng = np.random.default_rng(0)
    rows = np.unique(rng.integers(0, sd.num_faces, 20000))
    rows = np.concatenate([rows, [0, sd.num_faces - 1]])
    rows = np.unique(rows)
    dofs = sd.cell_faces.indices.reshape(sd.num_cells, 4)
    ref = sampled_rows_reference("rt0", "mass", sd, rows, dofs)
    ip = Ad.indptr.cpu().numpy()
    rt = torch.from_numpy(rows).cuda()
    lo, hi = ip[rows], ip[rows + 1]
    assert np.array_equal(hi - lo, ref.indptr[rows + 1] - ref.indptr[rows])
    pos = torch.from_numpy(np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)])).cuda()
    cols, vals = Ad.indices[pos].cpu().numpy(), Ad.data[pos].cpu().numpy()
    rpos = np.concatenate([np.arange(a, b) for a, b in zip(ref.indptr[rows], ref.indptr[rows + 1])])
    assert np.array_equal(cols, ref.indices[rpos])
    assert np.abs(vals - ref.data[rpos]).max() <= TOL * np.abs(ref.data).max()
    del rt
