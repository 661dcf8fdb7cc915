"""Block-matrix helpers for arrays of sparse blocks (``utils/bmat.py:9-99``): the mixed-dimensional
operators are assembled as 2-D object arrays with one block per (subdomain, subdomain) pair."""

from __future__ import annotations

import numpy as np
import scipy.sparse as sps


def find_row_col_lengths(mat: np.ndarray):
    """Row heights and column widths implied by the blocks that are present."""
    rows = np.zeros(mat.shape[0], dtype=int)
    cols = np.zeros(mat.shape[1], dtype=int)
    for i in range(mat.shape[0]):
        for j in range(mat.shape[1]):
            if mat[i, j] is not None:
                rows[i], cols[j] = mat[i, j].shape
    return rows, cols


def replace_nones_with_zeros(mat: np.ndarray) -> None:
    """In place: every missing block becomes an empty sparse block of the right shape."""
    missing = [(i, j) for i in range(mat.shape[0]) for j in range(mat.shape[1]) if mat[i, j] is None]
    if not missing:
        return
    rows, cols = find_row_col_lengths(mat)
    for i, j in missing:
        mat[i, j] = sps.csc_array((rows[i], cols[j]))


def transpose(mat: np.ndarray) -> np.ndarray:
    """Block transpose that also transposes the blocks."""
    out = np.empty(mat.shape[::-1], dtype=object)
    for i in range(mat.shape[0]):
        for j in range(mat.shape[1]):
            out[j, i] = None if mat[i, j] is None else mat[i, j].T
    return out


def multiply(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Block product ``C[i, j] = sum_k A[i, k] @ B[k, j]``."""
    if A.shape[1] != B.shape[0]:
        raise ValueError("block shapes do not match")
    C = np.empty((A.shape[0], B.shape[1]), dtype=object)
    for i in range(A.shape[0]):
        for j in range(B.shape[1]):
            acc = A[i, 0] @ B[0, j]
            for k in range(1, A.shape[1]):
                acc = acc + A[i, k] @ B[k, j]
            C[i, j] = acc
    return C


def empty(n: int) -> np.ndarray:
    return np.empty((n, n), dtype=object)


def diag_blocks(blocks) -> np.ndarray:
    out = empty(len(blocks))
    for i, b in enumerate(blocks):
        out[i, i] = b
    return out
