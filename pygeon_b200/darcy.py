"""One-pass assembly of the mixed-Darcy RT0-P0 block system.

The reference stacks the blocks on the host with ``sps.block_array([[M, -B.T], [B, None]])``
(tutorials/fluid_flow/darcy.ipynb cell 10; tests/patch_tests/test_laplacian_mixed.py:50), where
``M = RT0.assemble_mass_matrix`` (fem/hdiv.py:255-274) and ``B = PwConstants mass @ div``
(fem/l2.py:335-353, fem/hdiv.py:177-192).  That matrix is not symmetric; MINRES needs the second
block row negated.  Here the **symmetrised** system

    [[ M, -B^T ],
     [ -B,  0  ]]      unknowns (face fluxes, cell pressures), rhs (f_u, -f_p)

is assembled directly on the device as one CSR: every cell contributes a (d+2)x(d+2) local block
over the dofs (its faces, num_faces + cell).  The zero pressure block is kept as explicit zeros on
the diagonal (structural pattern).
"""

from __future__ import annotations

import scipy.sparse as sps

from . import engine
from .constants import SECOND_ORDER_TENSOR, UNITARY_DATA, WEIGHT
from .params import coefficient


def darcy_saddle_device(sd, data: dict | None = None, keyword: str = UNITARY_DATA) -> engine.DeviceCSR:
    w = coefficient(sd, data, keyword, WEIGHT)
    K = coefficient(sd, data, keyword, SECOND_ORDER_TENSOR)
    return engine.assemble_device("darcy", "saddle", sd, w, K)


def darcy_saddle_matrix(sd, data: dict | None = None, keyword: str = UNITARY_DATA) -> sps.csc_array:
    """Host copy (CSC).  For non-symmetric tensors the CSR arrays are reinterpreted as CSC exactly as
    for the mass matrices (see pgb.h, K1 note)."""
    w = coefficient(sd, data, keyword, WEIGHT)
    K = coefficient(sd, data, keyword, SECOND_ORDER_TENSOR)
    return engine.assemble_host("darcy", "saddle", sd, w, K)


def darcy_block_jacobi(A: engine.DeviceCSR, num_faces: int):
    """Inverse diagonal of the SPD block preconditioner diag(diag(M), diag(B diag(M)^-1 B^T)) for
    the symmetrised saddle system (the natural scaling fix for SURVEY.md H5: ``B`` has entries
    +-1/|c|, so unpreconditioned MINRES stagnates).  One pass over the matrix in HBM
    (``pgb_darcy_block_jacobi``), scratch = one vector."""
    import torch

    from ._lib import check, lib

    n = A.shape[0]
    dev = A.data.device
    with torch.cuda.device(dev):
        work = torch.empty(num_faces, dtype=torch.float64, device=dev)
        dinv = torch.empty(n, dtype=torch.float64, device=dev)
        check(lib.pgb_darcy_block_jacobi(n, num_faces, A.indptr.data_ptr(), A.idx64, A.indices.data_ptr(),
                                         A.data.data_ptr(), work.data_ptr(), dinv.data_ptr(), engine._stream()))
    return dinv
