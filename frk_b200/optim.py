"""stats::optim(method = "Nelder-Mead") and stats::uniroot for the M-step's host-side control flow.

In the reference these optimisers run in R on the host and only their objective functions touch
the data (R/SREfit.R:386-388 uniroot on J(sigma2fs); :618-621 optim on f_tau).  The split is kept:
the iterate logic below is plain Python, every objective evaluation is a set of libfrkb200 kernels.
Ported from the published algorithms (R src/appl/optim.c:nmmin, R src/library/stats/src/zeroin.c)
so that the iterate path -- and therefore the log-likelihood trace -- follows the reference's.
"""
from __future__ import annotations

import numpy as np


def nmmin(fn, par, maxit=100, reltol=1e-4, abstol=-np.inf, alpha=1.0, bet=0.5, gamm=2.0):
    """R's optim(method="Nelder-Mead") core (src/appl/optim.c:nmmin), as called at
    R/SREfit.R:618-621 with control=list(maxit=100, reltol=1e-4)."""
    big = 1.0e35
    B = np.array(par, dtype=np.float64).copy()
    n = len(B)
    if maxit <= 0:
        return B, fn(B), 0, 0
    P = np.zeros((n + 1, n + 2))
    f = fn(B)
    if not np.isfinite(f):
        raise RuntimeError("function cannot be evaluated at initial parameters")
    funcount = 1
    convtol = reltol * (abs(f) + reltol)
    n1 = n + 1
    C = n + 2
    P[n1 - 1, 0] = f
    P[:n, 0] = B
    L = 1
    size = 0.0
    step = 0.0
    for i in range(n):
        if 0.1 * abs(B[i]) > step:
            step = 0.1 * abs(B[i])
    if step == 0.0:
        step = 0.1
    for j in range(2, n1 + 1):
        P[:n, j - 1] = B
        trystep = step
        while P[j - 2, j - 1] == B[j - 2]:
            P[j - 2, j - 1] = B[j - 2] + trystep
            trystep *= 10
        size += trystep
    oldsize = size
    calcvert = True
    fail = 0
    while True:
        if calcvert:
            for j in range(n1):
                if j + 1 != L:
                    B = P[:n, j].copy()
                    f = fn(B)
                    if not np.isfinite(f):
                        f = big
                    funcount += 1
                    P[n1 - 1, j] = f
            calcvert = False
        VL = P[n1 - 1, L - 1]
        VH = VL
        H = L
        for j in range(1, n1 + 1):
            if j != L:
                f = P[n1 - 1, j - 1]
                if f < VL:
                    L = j
                    VL = f
                if f > VH:
                    H = j
                    VH = f
        if VH <= VL + convtol or VL <= abstol:
            break
        for i in range(n):
            temp = -P[i, H - 1]
            for j in range(n1):
                temp += P[i, j]
            P[i, C - 1] = temp / n
        B = (1.0 + alpha) * P[:n, C - 1] - alpha * P[:n, H - 1]
        f = fn(B)
        if not np.isfinite(f):
            f = big
        funcount += 1
        VR = f
        if VR < VL:
            P[n1 - 1, C - 1] = f
            for i in range(n):
                f = gamm * B[i] + (1 - gamm) * P[i, C - 1]
                P[i, C - 1] = B[i]
                B[i] = f
            f = fn(B)
            if not np.isfinite(f):
                f = big
            funcount += 1
            if f < VR:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = f
            else:
                P[:n, H - 1] = P[:n, C - 1]
                P[n1 - 1, H - 1] = VR
        else:
            if VR < VH:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = VR
            B = (1 - bet) * P[:n, H - 1] + bet * P[:n, C - 1]
            f = fn(B)
            if not np.isfinite(f):
                f = big
            funcount += 1
            if f < P[n1 - 1, H - 1]:
                P[:n, H - 1] = B
                P[n1 - 1, H - 1] = f
            elif VR >= VH:
                calcvert = True
                size = 0.0
                for j in range(n1):
                    if j + 1 != L:
                        for i in range(n):
                            P[i, j] = bet * (P[i, j] - P[i, L - 1]) + P[i, L - 1]
                            size += abs(P[i, j] - P[i, L - 1])
                if size < oldsize:
                    oldsize = size
                else:
                    fail = 10
                    break
        if not (funcount <= maxit):
            break
    fmin = P[n1 - 1, L - 1]
    X = P[:n, L - 1].copy()
    if funcount > maxit:
        fail = 1
    return X, fmin, funcount, fail


def zeroin(f, lower, upper, tol=np.finfo(np.float64).eps ** 0.25, maxiter=1000):
    """stats::uniroot -> R_zeroin2 (Brent), default tol=.Machine$double.eps^0.25, as called
    at R/SREfit.R:386-388.  Returns the root."""
    EPS = np.finfo(np.float64).eps
    fa = f(lower)
    fb = f(upper)
    xmax = np.finfo(np.float64).max
    fa = min(max(fa, -xmax), xmax)
    fb = min(max(fb, -xmax), xmax)
    if fa * fb > 0:
        raise RuntimeError("f() values at end points not of opposite sign")
    a, b = lower, upper
    c, fc = a, fa
    if fa == 0.0:
        return a
    if fb == 0.0:
        return b
    maxit = maxiter + 1
    while maxit > 0:
        maxit -= 1
        prev_step = b - a
        if abs(fc) < abs(fb):
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol_act = 2 * EPS * abs(b) + tol / 2
        new_step = (c - b) / 2
        if abs(new_step) <= tol_act or fb == 0.0:
            return b
        if abs(prev_step) >= tol_act and abs(fa) > abs(fb):
            cb = c - b
            if a == c:
                t1 = fb / fa
                p = cb * t1
                q = 1.0 - t1
            else:
                q = fa / fc
                t1 = fb / fc
                t2 = fb / fa
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0))
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0)
            if p > 0:
                q = -q
            else:
                p = -p
            if p < (0.75 * cb * q - abs(tol_act * q) / 2) and p < abs(prev_step * q / 2):
                new_step = p / q
        if abs(new_step) < tol_act:
            new_step = tol_act if new_step > 0 else -tol_act
        a, fa = b, fb
        b += new_step
        fb = f(b)
        if (fb > 0 and fc > 0) or (fb < 0 and fc < 0):
            c, fc = a, fa
    return b
