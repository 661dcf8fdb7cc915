"""GPU Krylov solvers with the ``solver(A, b) -> x`` shape of ``pg.LinearSystem.solve``
(``numerics/linear_system.py:79-94``).

The recurrences and stopping rules are those of ``scipy.sparse.linalg.cg`` / ``minres``
(scipy 1.18.1), the oracle for this part; the iteration runs entirely on the device
(``pgb_cg`` / ``pgb_minres`` in libpgb.so).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .engine import DeviceCSR, _device, _ptr, _stream, _to_dev


@dataclass
class SolveInfo:
    iterations: int
    info: int
    residuals: np.ndarray


def _as_device_csr(A, device=None) -> DeviceCSR:
    if isinstance(A, DeviceCSR):
        return A
    return DeviceCSR.from_scipy(A, device)


def _vec(b, dev) -> torch.Tensor:
    if isinstance(b, torch.Tensor):
        return b.to(dev, torch.float64).contiguous()
    return _to_dev(np.ascontiguousarray(b, dtype=np.float64), dev)


RESID_CAP = 1 << 20  # PGB_RESID_CAP: residual-history entries kept per solve
_INT_MAX = 2**31 - 1


def _maxiter(maxiter, default: int) -> int:
    """scipy's default (``10 n`` for cg, ``5 n`` for minres) clamped to the C ABI's ``int``; an explicit
    value outside that range is an error, not a silent wrap-around."""
    if maxiter is None:
        return min(int(default), _INT_MAX)
    maxiter = int(maxiter)
    if not 0 <= maxiter <= _INT_MAX:
        raise ValueError(f"maxiter = {maxiter} does not fit the C ABI's int")
    return maxiter


def cg_device(A: DeviceCSR, b: torch.Tensor, x0=None, *, rtol=1e-5, atol=0.0, maxiter=None, dinv=None):
    """CG on device tensors.  Returns ``(x, SolveInfo)``; ``info`` is 0 or ``maxiter`` like scipy.
    ``SolveInfo.residuals`` holds at most ``RESID_CAP`` entries (the first ones)."""
    n = A.shape[0]
    dev = b.device
    maxiter = _maxiter(maxiter, 10 * n)
    with torch.cuda.device(dev):
        x = torch.zeros(n, dtype=torch.float64, device=dev) if x0 is None else x0.clone()
        work = torch.empty(3 * n, dtype=torch.float64, device=dev)
        resid = np.zeros(min(maxiter + 1, RESID_CAP))
        it, info = C.c_int(0), C.c_int(0)
        check(lib.pgb_cg(n, A.indptr.data_ptr(), A.idx64, A.indices.data_ptr(), A.data.data_ptr(),
                         b.data_ptr(), x.data_ptr(), _ptr(dinv), float(rtol), float(atol), int(maxiter),
                         work.data_ptr(), C.byref(it), C.byref(info), resid.ctypes.data, _stream()))
        _lib.count(2 + 3 * it.value)
    k = it.value
    return x, SolveInfo(k, info.value, resid[: min(k + 1, resid.size)].copy())


def minres_device(A: DeviceCSR, b: torch.Tensor, x0=None, *, rtol=1e-5, shift=0.0, maxiter=None, dinv=None):
    n = A.shape[0]
    dev = b.device
    maxiter = _maxiter(maxiter, 5 * n)
    with torch.cuda.device(dev):
        x = torch.zeros(n, dtype=torch.float64, device=dev) if x0 is None else x0.clone()
        work = torch.empty(6 * n, dtype=torch.float64, device=dev)
        resid = np.zeros(min(maxiter + 1, RESID_CAP))
        it, info = C.c_int(0), C.c_int(0)
        check(lib.pgb_minres(n, A.indptr.data_ptr(), A.idx64, A.indices.data_ptr(), A.data.data_ptr(),
                             b.data_ptr(), x.data_ptr(), _ptr(dinv), float(shift), float(rtol), int(maxiter),
                             work.data_ptr(), C.byref(it), C.byref(info), resid.ctypes.data, _stream()))
        _lib.count(5 + 4 * it.value)
    k = it.value
    return x, SolveInfo(k, info.value, resid[: min(k, resid.size)].copy())


def _columns(fn, A, b, device):
    """Apply a single-RHS device solver to ``b`` of shape (n,) or (n, k) (LinearSystem supports both,
    ``linear_system.py:96-112``)."""
    dev = _device(device)
    Ad = _as_device_csr(A, dev)
    b = np.asarray(b.todense()) if hasattr(b, "todense") else np.asarray(b)
    if b.ndim == 1:
        x, info = fn(Ad, _vec(b, dev))
        return x.cpu().numpy(), info
    outs, infos = [], []
    for j in range(b.shape[1]):
        x, info = fn(Ad, _vec(b[:, j], dev))
        outs.append(x.cpu().numpy())
        infos.append(info)
    return np.stack(outs, axis=1), infos


def cg(A, b, x0=None, *, rtol=1e-5, atol=0.0, maxiter=None, M=None, device=None):
    """Drop-in for ``scipy.sparse.linalg.cg(A, b, ...)`` -> ``(x, info)``.  ``M`` may be a 1-D array
    holding the inverse diagonal (Jacobi)."""
    dev = _device(device)
    dinv = None if M is None else _vec(np.asarray(M), dev)
    x0d = None if x0 is None else _vec(x0, dev)
    x, info = _columns(lambda Ad, bd: cg_device(Ad, bd, x0d, rtol=rtol, atol=atol, maxiter=maxiter, dinv=dinv),
                       A, b, dev)
    cg.last_info = info
    return x, (info.info if isinstance(info, SolveInfo) else max(i.info for i in info))


def minres(A, b, x0=None, *, rtol=1e-5, shift=0.0, maxiter=None, M=None, device=None):
    """Drop-in for ``scipy.sparse.linalg.minres(A, b, ...)`` -> ``(x, info)``.  ``M`` may be a 1-D
    array / tensor holding a positive inverse diagonal (Jacobi-type preconditioner)."""
    dev = _device(device)
    x0d = None if x0 is None else _vec(x0, dev)
    dinv = None if M is None else _vec(M if isinstance(M, torch.Tensor) else np.asarray(M), dev)
    x, info = _columns(lambda Ad, bd: minres_device(Ad, bd, x0d, rtol=rtol, shift=shift, maxiter=maxiter, dinv=dinv),
                       A, b, dev)
    minres.last_info = info
    return x, (info.info if isinstance(info, SolveInfo) else max(i.info for i in info))


def _make_solver(kind: str, host_fn, dev_fn, kw: dict):
    """``solver`` callable for ``LinearSystem.solve``.  Called as ``solver(A, b)`` it behaves like any
    scipy-style solver; ``pg.LinearSystem`` recognises ``pgb_kind`` and keeps the elimination of the
    essential dofs and the solve on the device (``linear_system.py:79-94`` is the host version)."""
    device = kw.pop("device", None)

    def solver(A, b):
        x, _ = host_fn(A, b, device=device, **kw)
        solver.info = host_fn.last_info
        return x

    def pgb_solve(A0, b0, keep, dls):
        dkw = dict(kw)
        M = dkw.pop("M", None)
        if M is not None:  # inverse diagonal over all dofs or over the free ones
            M = _vec(M if isinstance(M, torch.Tensor) else np.asarray(M), b0.device)
            dkw["dinv"] = M[keep.long()] if M.numel() != b0.numel() else M
        x0, info = dev_fn(A0, b0, **dkw)
        x = dls.ess_vals.clone()
        with torch.cuda.device(b0.device):
            check(lib.pgb_scatter_add(int(keep.numel()), keep.data_ptr(), x0.data_ptr(), x.data_ptr(), _stream()))
            _lib.count(1)
        return x, info

    solver.pgb_kind = kind
    solver.pgb_device = device
    solver.pgb_solve = pgb_solve
    solver.info = None
    return solver


def cg_solver(**kw):
    """``ls.solve(pg.cg_solver(rtol=1e-10))``: CG with scipy's recurrence and stopping rule."""
    return _make_solver("cg", cg, cg_device, dict(kw))


def minres_solver(**kw):
    return _make_solver("minres", minres, minres_device, dict(kw))
