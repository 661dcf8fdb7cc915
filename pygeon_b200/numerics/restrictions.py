"""Tip-dof restrictions (``numerics/restrictions.py:11-86``)."""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps

from .. import bmat

_ENTITY = ("cells", "faces", "ridges", "peaks")


def get_codim_str(n_minus_k: int) -> str:
    return _ENTITY[n_minus_k]


def zero_tip_dofs(mdg, n_minus_k: int, **kwargs):
    """Diagonal 0/1 operator that sends the dofs on tip entities to zero."""
    if n_minus_k == 0:
        return sps.diags_array(np.ones(mdg.num_subdomain_cells()), dtype=int).tocsc()
    tag = "tip_" + get_codim_str(n_minus_k)
    sds = mdg.subdomains()
    blocks = bmat.empty(len(sds))
    for i, sd in enumerate(sds):
        if sd.dim >= n_minus_k:
            blocks[i, i] = sps.diags_array(np.logical_not(sd.tags[tag]), dtype=int)
    bmat.replace_nones_with_zeros(blocks)
    return blocks if kwargs.get("as_bmat", False) else sps.block_array(blocks).tocsc()


def remove_tip_dofs(mdg, n_minus_k: int) -> sps.csc_array:
    """Rows of :func:`zero_tip_dofs` at the non-tip dofs (a restriction matrix)."""
    R = sps.csr_array(zero_tip_dofs(mdg, n_minus_k))
    return R[R.indices, :].tocsc()
