"""CPU oracle for the Krylov part.  TEST INFRASTRUCTURE ONLY.

The reference never calls an iterative solver: ``LinearSystem.solve(solver=spsolve)``
(``src/pygeon/numerics/linear_system.py:79-94``) takes any ``solver(A, b) -> x``.  The oracle is
therefore the third-party ``scipy.sparse.linalg.cg`` / ``minres`` (scipy 1.18.1, installed in the
image; ``_isolve/iterative.py`` and ``_isolve/minres.py``), restated here in plain numpy so that
the residual history is observable.  **Parity of Krylov iterates is unpinned by the reference's
own tests**; the pins are (i) agreement of these restatements with scipy itself
(tests/test_krylov_oracle.py) and (ii) the reference's patch tests' exact solutions.
"""

from __future__ import annotations

import numpy as np


def cg(A, b, x0=None, *, rtol=1e-5, atol=0.0, maxiter=None, dinv=None):
    """scipy.sparse.linalg.cg restated; returns (x, info, iterations, residual_norms)."""
    b = np.asarray(b, dtype=float)
    n = b.size
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=float)
    bnrm2 = np.linalg.norm(b)
    atol = max(float(atol), float(rtol) * float(bnrm2))
    if bnrm2 == 0:
        return b.copy(), 0, 0, np.zeros(0)
    if maxiter is None:
        maxiter = n * 10
    r = b - A @ x if x.any() else b.copy()
    hist = []
    rho_prev, p = None, None
    for it in range(maxiter):
        rn = np.linalg.norm(r)
        hist.append(rn)
        if rn < atol:
            return x, 0, it, np.array(hist)
        z = r if dinv is None else dinv * r
        rho_cur = np.dot(r, z)
        if it > 0:
            beta = rho_cur / rho_prev
            p *= beta
            p += z
        else:
            p = z.copy()
        q = A @ p
        alpha = rho_cur / np.dot(p, q)
        x += alpha * p
        r -= alpha * q
        rho_prev = rho_cur
    hist.append(np.linalg.norm(r))
    return x, maxiter, maxiter, np.array(hist)


def minres(A, b, x0=None, *, rtol=1e-5, shift=0.0, maxiter=None, dinv=None):
    """scipy.sparse.linalg.minres restated (optional inverse-diagonal preconditioner); returns
    (x, info, iterations, phibar history)."""
    b = np.asarray(b, dtype=float)
    n = b.size
    if maxiter is None:
        maxiter = 5 * n
    eps = np.finfo(float).eps
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=float)
    psolve = (lambda r: r) if dinv is None else (lambda r: dinv * r)
    r1 = b.copy() if x0 is None else b - A @ x
    y = psolve(r1)
    beta1 = np.dot(r1, y)
    if beta1 == 0:
        return x, 0, 0, np.zeros(0)
    if np.linalg.norm(b) == 0:
        return b.copy(), 0, 0, np.zeros(0)
    beta1 = np.sqrt(beta1)
    oldb, beta, dbar, epsln = 0.0, beta1, 0.0, 0.0
    phibar, rhs1, rhs2 = beta1, beta1, 0.0
    tnorm2, gmax, gmin = 0.0, 0.0, np.finfo(float).max
    cs, sn = -1.0, 0.0
    w = np.zeros(n)
    w2 = np.zeros(n)
    r2 = r1
    istop, itn = 0, 0
    hist = []
    while itn < maxiter:
        itn += 1
        v = (1.0 / beta) * y
        y = A @ v - shift * v
        if itn >= 2:
            y = y - (beta / oldb) * r1
        alfa = np.dot(v, y)
        y = y - (alfa / beta) * r2
        r1, r2 = r2, y
        y = psolve(r2)
        oldb = beta
        beta = np.sqrt(np.dot(r2, y))
        tnorm2 += alfa**2 + oldb**2 + beta**2
        if itn == 1 and beta / beta1 <= 10 * eps:
            istop = -1
        oldeps = epsln
        delta = cs * dbar + sn * alfa
        gbar = sn * dbar - cs * alfa
        epsln = sn * beta
        dbar = -cs * beta
        root = np.hypot(gbar, dbar)
        gamma = max(np.hypot(gbar, beta), eps)
        cs, sn = gbar / gamma, beta / gamma
        phi = cs * phibar
        phibar = sn * phibar
        w1, w2 = w2, w
        w = (v - oldeps * w1 - delta * w2) * (1.0 / gamma)
        x = x + phi * w
        gmax, gmin = max(gmax, gamma), min(gmin, gamma)
        z = rhs1 / gamma
        rhs1 = rhs2 - delta * z
        rhs2 = -epsln * z
        Anorm = np.sqrt(tnorm2)
        ynorm = np.linalg.norm(x)
        epsx = Anorm * ynorm * eps
        rnorm = phibar
        hist.append(phibar)
        test1 = np.inf if (ynorm == 0 or Anorm == 0) else rnorm / (Anorm * ynorm)
        test2 = np.inf if Anorm == 0 else root / Anorm
        Acond = gmax / gmin
        if istop == 0:
            if 1 + test2 <= 1:
                istop = 2
            if 1 + test1 <= 1:
                istop = 1
            if itn >= maxiter:
                istop = 6
            if Acond >= 0.1 / eps:
                istop = 4
            if epsx >= beta1:
                istop = 3
            if test2 <= rtol:
                istop = 2
            if test1 <= rtol:
                istop = 1
        if istop != 0:
            break
    return x, (maxiter if istop == 6 else 0), itn, np.array(hist)
