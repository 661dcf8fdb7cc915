"""``pg.cell_stiff / face_stiff / ridge_stiff / peak_stiff`` (``numerics/stiffness.py:10-109``):
``d^T M_{k+1} d`` with the exterior derivative ``d`` and the mass matrix of the next space."""

from __future__ import annotations

import scipy.sparse as sps

from . import differentials, innerproducts


def stiff_matrix(mdg, n_minus_k: int, discr, **kwargs) -> sps.csc_array:
    if n_minus_k == 0:
        n = mdg.num_subdomain_cells()
        return sps.csc_array((n, n))
    d = differentials.exterior_derivative(mdg, n_minus_k)
    mass = innerproducts.mass_matrix(mdg, n_minus_k - 1, discr, **kwargs)
    return (d.T @ mass @ d).tocsc()


def _alias(n_minus_k: int, name: str):
    def f(mdg, discr=None, **kwargs):
        return stiff_matrix(mdg, n_minus_k, discr, **kwargs)

    f.__name__ = f.__qualname__ = name
    f.__doc__ = f"Stiffness matrix of the space on the {name.split('_')[0]}s of a (mixed-dimensional) grid."
    return f


cell_stiff, face_stiff, ridge_stiff, peak_stiff = (
    _alias(k, n + "_stiff") for k, n in enumerate(("cell", "face", "ridge", "peak")))
