"""``pg.LinearSystem`` (``src/pygeon/numerics/linear_system.py:9-127``).

Holds ``A, b`` and the essential boundary conditions; ``solve(solver)`` eliminates the essential
dofs and hands the reduced system to ``solver(A_csc, b) -> x``, the plug-in point for the GPU
Krylov solvers in :mod:`pygeon_b200.solvers`.
"""

from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
import scipy.sparse as sps


def create_restriction(keep_dof: np.ndarray) -> sps.csc_array:
    """Rows of the identity at the kept dofs (``linear_system.py:115-127``)."""
    keep = np.flatnonzero(keep_dof)
    n = keep_dof.size
    return sps.csc_array(
        (np.ones(keep.size, dtype=int), (np.arange(keep.size), keep)), shape=(keep.size, n)
    )


class LinearSystem:
    def __init__(self, A, b: np.ndarray | None = None) -> None:
        self.A = A
        if b is None:
            b = np.zeros(A.shape[0])
        self.b = b
        self.reset_bc()

    def reset_bc(self) -> None:
        self.is_dof = np.ones(self.b.shape[0], dtype=bool)
        self.ess_vals = np.zeros(self.b.shape[0])

    def flag_ess_bc(self, is_ess_dof: np.ndarray, ess_vals: np.ndarray) -> None:
        self.is_dof[is_ess_dof] = False
        self.ess_vals[is_ess_dof] += ess_vals[is_ess_dof]

    def reduce_system(self) -> Tuple[sps.csc_array, np.ndarray, sps.csc_array]:
        R_0 = create_restriction(self.is_dof)
        A_0 = R_0 @ self.A @ R_0.T
        b_0 = R_0 @ (self.b - self.A @ self.repeat_ess_vals())
        return A_0, b_0, R_0

    def solve(self, solver: Callable | None = None) -> np.ndarray:
        if solver is None:
            solver = sps.linalg.spsolve
        A_0, b_0, R_0 = self.reduce_system()
        sol_0 = solver(A_0.tocsc(), b_0)
        return self.repeat_ess_vals() + R_0.T @ sol_0

    def repeat_ess_vals(self):
        if self.b.ndim == 1:
            return self.ess_vals
        elif not np.any(self.ess_vals):
            return sps.csc_array(self.b.shape)
        else:
            ess_vals = sps.csr_array(np.atleast_2d(self.ess_vals))
            return sps.vstack([ess_vals] * self.b.shape[1]).T.tocsc()
