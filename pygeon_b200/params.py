"""Coefficient convention of the reference (``params/data.py:10-118``).

``data = {PARAMETERS: {keyword: {SECOND_ORDER_TENSOR: sot, WEIGHT: w}}}``; missing entries
default to identity / 1, scalars are broadcast to all cells.  ``sot`` is anything with a
``.values`` array of shape (3, 3, num_cells) (``pp.SecondOrderTensor``) or a scalar.
"""

from __future__ import annotations

import numpy as np

from .constants import (AMBIENT_DIM, MATRIX, NORMAL_DIFFUSIVITY, PARAMETERS, SCALAR, SECOND_ORDER_TENSOR, VECTOR,
                        VECTOR_FIELD, WEIGHT)

data_default = {SECOND_ORDER_TENSOR: 1, VECTOR_FIELD: np.zeros(AMBIENT_DIM), WEIGHT: 1, NORMAL_DIFFUSIVITY: 1}


class SecondOrderTensor:
    """Symmetric 3x3 tensor per cell with ``.values`` (3, 3, nc); mirrors ``pp.SecondOrderTensor``."""

    def __init__(self, kxx, kyy=None, kzz=None, kxy=None, kxz=None, kyz=None):
        kxx = np.atleast_1d(np.asarray(kxx, dtype=float))
        z = np.zeros(kxx.size)
        kyy = kxx if kyy is None else np.asarray(kyy, dtype=float)
        kzz = kxx if kzz is None else np.asarray(kzz, dtype=float)
        kxy = z if kxy is None else np.asarray(kxy, dtype=float)
        kxz = z if kxz is None else np.asarray(kxz, dtype=float)
        kyz = z if kyz is None else np.asarray(kyz, dtype=float)
        self.values = np.array([[kxx, kxy, kxz], [kxy, kyy, kyz], [kxz, kyz, kzz]], dtype=float)

    def copy(self):
        out = SecondOrderTensor(np.ones(self.values.shape[2]))
        out.values = self.values.copy()
        return out


def initialize_data(data: dict, keyword: str, specified_parameters: dict | None = None) -> dict:
    """``pp.initialize_data``: ``data[PARAMETERS][keyword].update(params)``."""
    data.setdefault(PARAMETERS, {}).setdefault(keyword, {}).update(specified_parameters or {})
    return data


def _lookup(data, keyword: str, param: str):
    if not data or PARAMETERS not in data or keyword not in data[PARAMETERS]:
        return data_default[param]
    return data[PARAMETERS][keyword].get(param, data_default[param])


def get_cell_data(sd, data, keyword: str, param: str, tensor_order: int = SCALAR):
    """``pg.get_cell_data`` (``params/data.py:61-118``): the parameter stored under
    ``data[PARAMETERS][keyword][param]`` or its default; scalars are broadcast to one value per cell
    (wrapped in a :class:`SecondOrderTensor` for ``tensor_order == MATRIX``), a single vector is
    repeated for every cell when ``tensor_order == VECTOR``.  ``sd`` may be a mortar grid."""
    value = _lookup(data, keyword, param)
    if isinstance(value, np.ScalarType):
        value = np.full(sd.num_cells, value)
        if tensor_order == MATRIX:
            value = SecondOrderTensor(value)
    if isinstance(value, np.ndarray) and tensor_order == VECTOR and value.ndim == 1:
        value = np.tile(value.reshape(-1, 1), (1, sd.num_cells))
    return value


def coefficient(sd, data, keyword: str, param: str):
    """The engine's view of a coefficient: ``None`` when it has its default value (nothing is
    uploaded, the kernels use 1 / identity), otherwise a validated float64 array.  Returns

    * for ``WEIGHT``: ``None`` (all ones) or a float64 array (nc,)
    * for ``SECOND_ORDER_TENSOR``: ``None`` (identity) or a float64 array (3, 3, nc)
    """
    value = _lookup(data, keyword, param)
    nc = sd.num_cells
    if param == WEIGHT:
        if np.isscalar(value):
            return None if value == 1 else np.full(nc, float(value))
        value = np.ascontiguousarray(value, dtype=np.float64)
        if value.shape != (nc,):
            raise ValueError(f"weight must have shape ({nc},)")
        return value
    if param == SECOND_ORDER_TENSOR:
        if np.isscalar(value):
            if value == 1:
                return None
            K = np.zeros((3, 3, nc))
            K[0, 0] = K[1, 1] = K[2, 2] = float(value)
            return K
        K = np.ascontiguousarray(value.values, dtype=np.float64)
        if K.shape != (3, 3, nc):
            raise ValueError(f"second order tensor must have shape (3, 3, {nc})")
        return K
    raise KeyError(param)
