"""``pg.cell_mass / face_mass / ridge_mass / peak_mass`` and their lumped variants
(``numerics/innerproducts.py:13-375``): per-subdomain dispatch to
``discr.assemble_mass_matrix(sd, data)`` (the CUDA path), block-diagonal stacking, plus the mortar
trace term of the face mass matrix."""

from __future__ import annotations

from typing import Callable

import numpy as np
import scipy.sparse as sps

from .. import bmat
from ..constants import NORMAL_DIFFUSIVITY, UNITARY_DATA
from ..params import get_cell_data


def default_discr(sd, n_minus_k: int, **kwargs):
    """Whitney forms: P0 (cells), RT0 (faces), Nedelec0 (ridges in 3D), Lagrange1 (peaks)."""
    from .. import discretization as D

    keyword = kwargs.get("keyword", UNITARY_DATA)
    if n_minus_k == 0:
        return D.PwConstants(keyword)
    if n_minus_k == 1:
        return D.RT0(keyword)
    if n_minus_k == sd.dim:
        return D.Lagrange1(keyword)
    if n_minus_k == 2:
        return D.Nedelec0(keyword)
    raise ValueError(f"no default discretization for n - k = {n_minus_k} in dimension {sd.dim}")


def _sd_matrix(method: str):
    def assemble(sd, n_minus_k: int, discr=None, data: dict | None = None, **kwargs):
        if n_minus_k > sd.dim:
            return sps.csc_array((0, 0))
        if discr is None:
            discr = default_discr(sd, n_minus_k, **kwargs)
        return getattr(discr, method)(sd, data)

    return assemble


_sd_mass_matrix = _sd_matrix("assemble_mass_matrix")
_sd_lumped_mass = _sd_matrix("assemble_lumped_matrix")


def local_matrix(sd, n_minus_k: int, discr, d_sd: dict, **kwargs):
    return _sd_mass_matrix(sd, n_minus_k, discr, d_sd, **kwargs)


def mass_matrix(mdg, n_minus_k: int, discr=None, local_matrix: Callable = local_matrix, **kwargs):
    """Block-diagonal mass matrix over the subdomains; for the face space the interfaces add
    ``S diag(1 / (|m| k_n)) S^T`` with ``S = signed_mortar_to_primary`` to the block of the
    higher-dimensional neighbour (``innerproducts.py:162-232``)."""
    if "keyword" in kwargs:
        keyword = kwargs["keyword"]
    elif discr is not None:
        keyword = discr.keyword
    else:
        keyword = UNITARY_DATA
    sds = mdg.subdomains(return_data=True)
    blocks = bmat.diag_blocks([local_matrix(sd, n_minus_k, discr, d_sd, **kwargs) for sd, d_sd in sds])
    if n_minus_k == 1 and kwargs.get("trace_contribution", True):
        index = {id(sd): i for i, (sd, _) in enumerate(sds)}
        for intf, d_intf in mdg.interfaces(return_data=True):
            i = index[id(mdg.interface_to_subdomain_pair(intf)[0])]
            kn = get_cell_data(intf, d_intf, keyword, NORMAL_DIFFUSIVITY)
            S = intf.signed_mortar_to_primary
            blocks[i, i] = blocks[i, i] + S @ sps.diags_array(1.0 / intf.cell_volumes / kn) @ S.T
    bmat.replace_nones_with_zeros(blocks)
    return blocks if kwargs.get("as_bmat", False) else sps.block_array(blocks).tocsc()


def lumped_mass_matrix(mdg, n_minus_k: int, discr=None, **kwargs):
    return mass_matrix(mdg, n_minus_k, discr, _sd_lumped_mass, **kwargs)


def _alias(fn, n_minus_k: int, name: str, what: str):
    def f(mdg, discr=None, **kwargs):
        return fn(mdg, n_minus_k, discr, **kwargs)

    f.__name__ = f.__qualname__ = name
    f.__doc__ = f"{what} of the space on the {_NAMES[n_minus_k]} of a (mixed-dimensional) grid."
    return f


_NAMES = ("cells", "faces", "ridges", "peaks")
cell_mass, face_mass, ridge_mass, peak_mass = (
    _alias(mass_matrix, k, f"{n[:-1]}_mass", "Mass matrix") for k, n in enumerate(_NAMES))
lumped_cell_mass, lumped_face_mass, lumped_ridge_mass, lumped_peak_mass = (
    _alias(lumped_mass_matrix, k, f"lumped_{n[:-1]}_mass", "Lumped mass matrix") for k, n in enumerate(_NAMES))
