"""Device-resident counterpart of ``pg.LinearSystem``.

``LinearSystem.reduce_system`` (``numerics/linear_system.py:64-77``) eliminates the essential
dofs on the host with two scipy SpGEMMs (``R @ A @ R.T``); at 10^8 cells that forces a round trip
of a multi-GB matrix.  Here the same reduction runs where the matrix lives: ``pgb_mask_to_map``
numbers the free dofs, ``pgb_csr_extract_*`` select their rows and columns (count -> scan -> fill,
O(n) scratch), ``b_0 = R (b - A ess)`` is one SpMV + one gather, and the reduced system goes
straight to the GPU Krylov drivers.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .engine import DeviceCSR, _stream, _to_dev, mask_to_map
from .solvers import cg_device, minres_device


def _as_dev(x, dev, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dev, dtype)
    return _to_dev(np.ascontiguousarray(x), dev, dtype)


class DeviceLinearSystem:
    def __init__(self, A: DeviceCSR, b) -> None:
        self.A = A
        self.device = A.data.device
        n = A.shape[0]
        self.b = torch.zeros(n, dtype=torch.float64, device=self.device) if b is None else _as_dev(
            b, self.device, torch.float64)
        self.reset_bc()

    def reset_bc(self) -> None:
        n = self.A.shape[0]
        self.is_dof = torch.ones(n, dtype=torch.bool, device=self.device)
        self.ess_vals = torch.zeros(n, dtype=torch.float64, device=self.device)

    def flag_ess_bc(self, is_ess_dof, ess_vals) -> None:
        m = _as_dev(is_ess_dof, self.device, torch.bool)
        v = _as_dev(ess_vals, self.device, torch.float64)
        self.is_dof &= ~m
        self.ess_vals += torch.where(m, v, torch.zeros_like(v))

    def reduce_system(self):
        """(A_0, b_0, kept dof ids): rows / columns of the free dofs, right-hand side corrected by
        the essential values."""
        A, dev = self.A, self.device
        n = A.shape[0]
        newid, keep = mask_to_map(self.is_dof)
        n0 = int(keep.numel())
        y = A.matvec(self.ess_vals)  # A ess
        b0 = torch.empty(n0, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib.pgb_gather(n0, keep.data_ptr(), self.b.data_ptr(), y.data_ptr(), -1.0, b0.data_ptr(), _stream()))
            _lib.count(1)
        del y
        A0 = A.extract(keep, newid, n0)
        return A0, b0, keep

    def solve(self, solver: str = "cg", **kw):
        """Solve on the device; returns (x (device tensor, all dofs), SolveInfo)."""
        A0, b0, keep = self.reduce_system()
        fn = {"cg": cg_device, "minres": minres_device}[solver]
        x0, info = fn(A0, b0, **kw)
        x = self.ess_vals.clone()
        with torch.cuda.device(self.device):
            check(lib.pgb_scatter_add(int(keep.numel()), keep.data_ptr(), x0.data_ptr(), x.data_ptr(), _stream()))
            _lib.count(1)
        return x, info
