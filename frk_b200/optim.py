"""stats::optim(method = "Nelder-Mead") and stats::uniroot as the M-step uses them (R/SREfit.R:618-621, :386-388).

The implementations live in libfrkb200 (csrc/model.cu: restated from the published algorithms, R src/appl/optim.c nmmin and
R src/library/stats/src/zeroin.c) because the whole EM loop runs inside ``frk_model_em_fit``; these wrappers expose them
through the C ABI (``frk_optim_nelder_mead`` / ``frk_optim_zeroin``) so that the iterate path can be pinned on its own --
they are pure host code and need no GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import OBJECTIVE_FN, call, vp


def nmmin(fn, par, maxit=100, reltol=1e-4):
    """Returns (x, f(x), function count, fail code) like R's optim()$par, $value, $counts[1], $convergence."""
    par = np.ascontiguousarray(par, dtype=np.float64)
    n = len(par)
    cb = OBJECTIVE_FN(lambda user, k, x: float(fn(np.ctypeslib.as_array(x, shape=(k,)).copy())))
    x = np.zeros(n)
    f, cnt, fail = C.c_double(0), C.c_int(0), C.c_int(0)
    call("frk_optim_nelder_mead", cb, vp(None), n, vp(par.ctypes.data), int(maxit), float(reltol), vp(x.ctypes.data),
         C.byref(f), C.byref(cnt), C.byref(fail))
    return x, f.value, cnt.value, fail.value


def nmmin_speculative(fn, par, width, maxit=100, reltol=1e-4):
    """The same minimisation, driven as the block-exponential M-step drives it on the device: whenever Nelder-Mead asks for a new
    point, that point and up to ``width - 1`` points it may ask for next are evaluated in one round.  Returns
    (x, f(x), points requested, rounds, points evaluated); x and f(x) equal :func:`nmmin`'s."""
    par = np.ascontiguousarray(par, dtype=np.float64)
    n = len(par)
    cb = OBJECTIVE_FN(lambda user, k, x: float(fn(np.ctypeslib.as_array(x, shape=(k,)).copy())))
    x = np.zeros(n)
    f, req, rounds, evals = C.c_double(0), C.c_int(0), C.c_int(0), C.c_int(0)
    call("frk_optim_nelder_mead_speculative", cb, vp(None), n, vp(par.ctypes.data), int(maxit), float(reltol), int(width),
         vp(x.ctypes.data), C.byref(f), C.byref(req), C.byref(rounds), C.byref(evals))
    return x, f.value, req.value, rounds.value, evals.value


def zeroin(f, lower, upper):
    """stats::uniroot(f, c(lower, upper))$root with the default tol = .Machine$double.eps^0.25."""
    cb = OBJECTIVE_FN(lambda user, k, x: float(f(x[0])))
    root = C.c_double(0)
    call("frk_optim_zeroin", cb, vp(None), float(lower), float(upper), C.byref(root))
    return root.value
