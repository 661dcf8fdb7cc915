"""Mixed-dimensional RT0 / P0 Darcy system resident on the GPU (BASELINE config #5).

The reference builds the block operators on the host: ``pg.face_mass(mdg)`` (block diagonal of the
subdomains' RT0 mass matrices plus the mortar trace term, ``numerics/innerproducts.py:162-232``),
``pg.cell_mass(mdg) @ pg.div(mdg)`` (``numerics/differentials.py:144-192``: subdomain divergences, the jump
operator of every interface, tip dofs masked) and ``sps.block_array`` for the saddle stack.  Here the large
part never leaves HBM:

* every subdomain's mass matrix is a CUDA assembly (``assemble_mass_matrix_device``);
* the mortar term ``S diag(1 / (|m| k_n)) S^T`` of a conforming interface is diagonal on the fracture faces of
  the higher-dimensional grid and is added in place (``pgb_csr_add_diagonal``);
* ``B`` (four entries per cell: topology times a diagonal) comes from the host operators and is uploaded once,
  ``B^T`` is made by the radix-sort transpose (``pgb_csr_transpose``);
* the blocks are stacked by the count -> scan -> place kernels (``engine.device_bmat``).

The result is the symmetrised saddle matrix ``[[M, -B^T], [-B, 0]]`` as a ``DeviceCSR`` ready for
``pg.minres_device`` with ``pg.darcy_block_jacobi`` as the preconditioner."""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import torch

from . import engine
from .constants import NORMAL_DIFFUSIVITY, UNITARY_DATA
from .params import get_cell_data


def _trace_diagonal(intf, d_intf, keyword):
    """(rows, values) of the mortar trace term on the primary grid's faces; raises when the term is not
    diagonal (a non-conforming interface needs the general host path, ``pg.face_mass``)."""
    kn = get_cell_data(intf, d_intf, keyword, NORMAL_DIFFUSIVITY)
    S = sps.csr_array(intf.signed_mortar_to_primary)
    T = sps.csr_array(S @ sps.diags_array(1.0 / intf.cell_volumes / kn) @ S.T)
    T.eliminate_zeros()
    d = T.diagonal()
    if T.nnz != np.count_nonzero(d):
        raise NotImplementedError("the mortar trace term is not diagonal: use pg.face_mass (host stacking)")
    rows = np.flatnonzero(d)
    return rows, d[rows]


def face_mass_device(mdg, discr=None, keyword: str = UNITARY_DATA, trace_contribution: bool = True,
                     device=None) -> engine.DeviceCSR:
    """``pg.face_mass(mdg)`` as one device CSR (``innerproducts.py:162-232`` with ``n_minus_k = 1``)."""
    from .discretization import RT0

    keyword = discr.keyword if discr is not None else keyword
    sds = mdg.subdomains(return_data=True)
    index = {id(sd): i for i, (sd, _) in enumerate(sds)}
    mats = []
    for sd, d_sd in sds:
        disc = discr if discr is not None else RT0(keyword)
        if disc.ndof(sd) == 0:
            raise NotImplementedError("face_mass_device: a subdomain without face dofs (dimension 0)")
        engine.get_pack(sd, device)
        mats.append(disc.assemble_mass_matrix_device(sd, d_sd))
    if trace_contribution:
        for intf, d_intf in mdg.interfaces(return_data=True):
            A = mats[index[id(mdg.interface_to_subdomain_pair(intf)[0])]]
            rows, vals = _trace_diagonal(intf, d_intf, keyword)
            A.add_diagonal(torch.from_numpy(rows.astype(np.int32)), torch.from_numpy(np.ascontiguousarray(vals)))
    n = len(mats)
    blocks = [[mats[i] if i == j else None for j in range(n)] for i in range(n)]
    return engine.device_bmat(blocks, symmetric=all(m.symmetric for m in mats))


def cell_div_device(mdg, keyword: str = UNITARY_DATA, device=None) -> engine.DeviceCSR:
    """``pg.cell_mass(mdg) @ pg.div(mdg)`` uploaded as a device CSR (topology scaled by 1 / |c|: built by the
    host operators, 4 entries per cell)."""
    from .numerics import differentials, innerproducts

    B = _left_scaled(innerproducts.cell_mass(mdg, keyword=keyword), differentials.div(mdg))
    return engine.DeviceCSR.from_scipy(B, device)


def _left_scaled(Cm, D) -> sps.csr_array:
    """``Cm @ D`` as CSR.  A diagonal ``Cm`` (the P0 mass matrix) scales the rows of ``D``: one pass over the
    entries instead of a sparse product, with the same values and - exact zeros dropped, as scipy's product does -
    the same pattern."""
    Cm = sps.csc_array(Cm)
    n = Cm.shape[0]
    if Cm.shape[0] == Cm.shape[1] and Cm.nnz == n and np.array_equal(Cm.indices, np.arange(n)):
        B = sps.csr_array(D, copy=True)
        B.data = B.data * np.repeat(Cm.data, np.diff(B.indptr))
        B.eliminate_zeros()
        return B
    return sps.csr_array(Cm @ D)


def darcy_saddle_device(mdg, keyword: str = UNITARY_DATA, discr=None, device=None):
    """``(A, num_face_dofs, num_cell_dofs)``: the symmetrised mixed-dimensional saddle matrix
    ``[[M, -B^T], [-B, 0]]`` on the device; unknowns (fluxes of all subdomains, pressures of all subdomains),
    right-hand side ``(f_u, -f_p)``.  The pressure block carries no stored entries."""
    M = face_mass_device(mdg, discr, keyword, device=device)
    B = cell_div_device(mdg, keyword, device=M.data.device)
    Bt = B.transpose()
    A = engine.device_bmat([[M, (Bt, -1.0)], [(B, -1.0), None]], symmetric=M.symmetric)
    return A, M.shape[0], B.shape[0]
