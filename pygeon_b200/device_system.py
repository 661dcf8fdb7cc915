"""Device-resident counterpart of ``pg.LinearSystem``.

``LinearSystem.reduce_system`` (``numerics/linear_system.py:64-77``) eliminates the essential
dofs on the host with two scipy SpGEMMs (``R @ A @ R.T``); at 10^8 cells that forces a round trip
of a multi-GB matrix.  Here the same reduction is done on the matrix where it lives (HBM): rows
and columns of the kept dofs are selected and renumbered, ``b_0 = R (b - A ess)`` uses the SpMV
kernel, and the reduced system goes straight to the GPU Krylov drivers.  torch is used for the
index bookkeeping (masks, prefix sums, compaction); the arithmetic (SpMV, CG/MINRES) is libpgb.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .engine import DeviceCSR, _stream, _to_dev
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
        self.is_dof[m] = False
        self.ess_vals[m] += v[m]

    def reduce_system(self):
        """(A_0, b_0, kept indices): rows/columns of the free dofs, rhs corrected by the essential
        values.  Valid for the symmetric-pattern matrices of this package (CSR == CSC pattern)."""
        A, dev = self.A, self.device
        n = A.shape[0]
        keep = torch.nonzero(self.is_dof).flatten()
        n0 = int(keep.numel())
        # b_0 = R (b - A ess)
        y = torch.empty(n, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib.pgb_spmv(n, A.indptr.data_ptr(), A.idx64, A.indices.data_ptr(), A.data.data_ptr(),
                               self.ess_vals.data_ptr(), y.data_ptr(), 0, _stream()))
            _lib.count(1)
        b0 = (self.b - y)[keep].contiguous()
        # select rows, then drop the eliminated columns
        newid = torch.full((n,), -1, dtype=torch.int64, device=dev)
        newid[keep] = torch.arange(n0, device=dev)
        ip = A.indptr.to(torch.int64)
        lens = ip[keep + 1] - ip[keep]
        row_of = torch.repeat_interleave(torch.arange(n0, device=dev), lens)
        start = torch.zeros(n0 + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens, 0, out=start[1:])
        pos = ip[keep][row_of] + (torch.arange(int(start[-1].item()), device=dev) - start[:-1][row_of])
        cols = newid[A.indices[pos].to(torch.int64)]
        ok = cols >= 0
        cols, vals, row_of = cols[ok], A.data[pos][ok], row_of[ok]
        counts = torch.bincount(row_of, minlength=n0)
        indptr = torch.zeros(n0 + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=indptr[1:])
        it = torch.int64 if int(indptr[-1].item()) >= 2**31 else torch.int32
        A0 = DeviceCSR((n0, n0), indptr.to(it), cols.to(torch.int32).contiguous(), vals.contiguous())
        return A0, b0, keep

    def solve(self, solver: str = "cg", **kw):
        """Solve on the device; returns (x (device tensor, all dofs), SolveInfo)."""
        A0, b0, keep = self.reduce_system()
        fn = {"cg": cg_device, "minres": minres_device}[solver]
        x0, info = fn(A0, b0, **kw)
        x = self.ess_vals.clone()
        x[keep] += x0
        return x, info
