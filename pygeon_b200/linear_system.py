"""``pg.LinearSystem`` (reference ``src/pygeon/numerics/linear_system.py:9-127``) as a front end of
the device path.

Same surface as the reference: ``LinearSystem(A, b)``, ``flag_ess_bc(is_ess_dof, ess_vals)``,
``reset_bc()``, ``reduce_system() -> (A_0, b_0, R_0)``, ``solve(solver) -> x`` with ``b`` of shape
``(n,)`` or ``(n, k)``.  What differs is where the work happens:

* ``solve(pg.cg_solver(...))`` / ``solve(pg.minres_solver(...))`` (callables made by
  :mod:`pygeon_b200.solvers`, recognised by their ``pgb_kind`` attribute) never builds the reduced
  system on the host: the matrix goes to HBM once (or already lives there as a
  :class:`~pygeon_b200.engine.DeviceCSR`), :class:`~pygeon_b200.device_system.DeviceLinearSystem`
  eliminates the essential dofs with the extraction kernels and the Krylov driver runs on the result;
* any other callable gets ``solver(A_0.tocsc(), b_0)`` like the reference, with ``A_0`` / ``b_0``
  obtained by index selection (no restriction-matrix products).
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import scipy.sparse as sps


def create_restriction(keep_dof: np.ndarray) -> sps.csc_array:
    """The rows of the identity that belong to the kept dofs (``linear_system.py:115-127``)."""
    cols = np.flatnonzero(keep_dof)
    m, n = cols.size, np.asarray(keep_dof).size
    return sps.csc_array((np.ones(m, dtype=int), (np.arange(m), cols)), shape=(m, n))


class LinearSystem:
    def __init__(self, A, b: np.ndarray | None = None) -> None:
        self.A = A
        self.b = np.zeros(A.shape[0]) if b is None else b
        self.reset_bc()

    # ---- essential boundary conditions ------------------------------------------------------
    def reset_bc(self) -> None:
        n = self.b.shape[0]
        self.is_dof = np.full(n, True)
        self.ess_vals = np.zeros(n, dtype=float)

    def flag_ess_bc(self, is_ess_dof: np.ndarray, ess_vals: np.ndarray) -> None:
        sel = np.flatnonzero(is_ess_dof) if np.asarray(is_ess_dof).dtype == bool else np.asarray(is_ess_dof)
        self.is_dof[sel] = False
        np.add.at(self.ess_vals, sel, np.asarray(ess_vals)[sel])

    def repeat_ess_vals(self):
        """Essential values shaped like ``b`` (one copy per right-hand side)."""
        if self.b.ndim == 1:
            return self.ess_vals
        return np.repeat(self.ess_vals[:, None], self.b.shape[1], axis=1)

    # ---- host elimination ---------------------------------------------------------------------
    def _host_matrix(self):
        A = self.A
        return A.to_scipy("csr") if hasattr(A, "to_scipy") else sps.csr_array(A)

    def reduce_system(self):
        """``(A_0, b_0, R_0)`` with ``A_0 = R_0 A R_0^T`` and ``b_0 = R_0 (b - A ess)``."""
        keep = np.flatnonzero(self.is_dof)
        A = self._host_matrix()
        lifted = self.b - A @ self.repeat_ess_vals()
        A_0 = A[keep][:, keep]
        return A_0, lifted[keep], create_restriction(self.is_dof)

    # ---- solve ------------------------------------------------------------------------------------
    def solve(self, solver: Callable | None = None) -> np.ndarray:
        if getattr(solver, "pgb_kind", None) is not None:
            return self._solve_device(solver)
        if solver is None:
            solver = sps.linalg.spsolve
        A_0, b_0, _ = self.reduce_system()
        sol_0 = solver(A_0.tocsc(), b_0)
        sol = np.array(self.repeat_ess_vals(), dtype=float, copy=True)
        sol[self.is_dof] += np.asarray(sol_0).reshape(sol[self.is_dof].shape)
        return sol

    def _solve_device(self, solver) -> np.ndarray:
        """Elimination + Krylov solve in HBM; only ``b`` goes up and ``x`` comes down."""
        from .device_system import DeviceLinearSystem
        from .engine import DeviceCSR

        A = self.A if isinstance(self.A, DeviceCSR) else DeviceCSR.from_scipy(self.A, solver.pgb_device)
        b = np.asarray(self.b.todense()) if hasattr(self.b, "todense") else np.asarray(self.b, dtype=float)
        cols = b[:, None] if b.ndim == 1 else b
        dls = DeviceLinearSystem(A, None)
        dls.flag_ess_bc(~self.is_dof, self.ess_vals)
        A0 = keep = None
        out, infos = [], []
        for j in range(cols.shape[1]):
            dls.b = dls.b.new_tensor(np.ascontiguousarray(cols[:, j]))
            if A0 is None:
                A0, b0, keep = dls.reduce_system()
            else:  # same matrix and constraints: only the right-hand side changes
                _, b0, _ = _rhs_only(dls, keep)
            x, info = solver.pgb_solve(A0, b0, keep, dls)
            out.append(x.cpu().numpy())
            infos.append(info)
        solver.info = infos[0] if b.ndim == 1 else infos
        return out[0] if b.ndim == 1 else np.stack(out, axis=1)


def _rhs_only(dls, keep):
    import torch

    from . import _lib
    from ._lib import check, lib
    from .engine import _stream

    y = dls.A.matvec(dls.ess_vals)
    b0 = torch.empty(int(keep.numel()), dtype=torch.float64, device=dls.device)
    with torch.cuda.device(dls.device):
        check(lib.pgb_gather(int(keep.numel()), keep.data_ptr(), dls.b.data_ptr(), y.data_ptr(), -1.0,
                             b0.data_ptr(), _stream()))
        _lib.count(1)
    return None, b0, keep
