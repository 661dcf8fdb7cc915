"""``pg.eval_at_cell_centers`` over the subdomains (``numerics/projections.py:11-49``)."""

from __future__ import annotations

import scipy.sparse as sps

from .. import bmat


def eval_at_cell_centers(mdg, discr, **kwargs):
    blocks = bmat.diag_blocks([discr.eval_at_cell_centers(sd) for sd in mdg.subdomains()])
    bmat.replace_nones_with_zeros(blocks)
    return blocks if kwargs.get("as_bmat", False) else sps.block_array(blocks).tocsc()
