"""``pg.div / pg.curl / pg.grad`` (``numerics/differentials.py:21-192``).

On a single grid (or a mortar grid) the exterior derivative of the (n-k)-forms is a transposed
incidence matrix of the grid; on a mixed-dimensional grid the diagonal blocks are the subdomains'
derivatives, the block (lower, upper) is the jump operator stored on the mortar grid, and tip
dofs are masked out.
"""

from __future__ import annotations

import warnings

import numpy as np
import scipy.sparse as sps

from .. import bmat
from ..md_grid import MixedDimensionalGrid, MortarGrid
from . import restrictions

_INCIDENCE = {1: "cell_faces", 2: "face_ridges", 3: "ridge_peaks"}


def _is_single_grid(grid) -> bool:
    return isinstance(grid, MortarGrid) or (hasattr(grid, "cell_faces") and hasattr(grid, "num_cells"))


def exterior_derivative(grid, n_minus_k: int, **kwargs):
    if isinstance(grid, MixedDimensionalGrid):
        return _mdg_exterior_derivative(grid, n_minus_k, **kwargs)
    if _is_single_grid(grid):
        return _g_exterior_derivative(grid, n_minus_k, **kwargs)
    raise TypeError("Input needs to be a grid, a mortar grid or a mixed-dimensional grid")


def div(grid, **kwargs):
    return exterior_derivative(grid, 1, **kwargs)


def curl(grid, **kwargs):
    return exterior_derivative(grid, 2, **kwargs)


def grad(grid, **kwargs):
    return exterior_derivative(grid, 3, **kwargs)


def _g_exterior_derivative(grid, n_minus_k: int, **_kwargs):
    """Single grid (``differentials.py:111-141``)."""
    if n_minus_k in _INCIDENCE:
        return getattr(grid, _INCIDENCE[n_minus_k]).T
    if n_minus_k == 0:
        return sps.csc_array((0, grid.num_cells))
    if n_minus_k == 4:
        return sps.csc_array((grid.num_peaks, 0))
    warnings.warn("(n - k) is not between 0 and 4")
    return sps.csc_array((0, 0))


def _mdg_exterior_derivative(mdg: MixedDimensionalGrid, n_minus_k: int, **kwargs):
    """Mixed-dimensional (``differentials.py:144-192``)."""
    sds = mdg.subdomains()
    blocks = bmat.diag_blocks([exterior_derivative(sd, n_minus_k) for sd in sds])
    index = {id(sd): i for i, sd in enumerate(sds)}
    for intf in mdg.interfaces():
        up, down = mdg.interface_to_subdomain_pair(intf)
        if up.dim >= n_minus_k:
            blocks[index[id(down)], index[id(up)]] = exterior_derivative(intf, n_minus_k)
    bmat.replace_nones_with_zeros(blocks)
    as_bmat = kwargs.get("as_bmat", False)
    no_tips = restrictions.zero_tip_dofs(mdg, n_minus_k, **kwargs)
    if as_bmat:
        return bmat.multiply(blocks, no_tips)
    return sps.block_array(blocks) @ no_tips
